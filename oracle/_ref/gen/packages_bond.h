#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "bond_fene.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "bond_fene_expand.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "bond_gromos.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "bond_harmonic.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "bond_morse.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "bond_quartic.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "bond_table.h"
