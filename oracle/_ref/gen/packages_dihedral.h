#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "dihedral_charmm.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "dihedral_charmmfsw.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "dihedral_harmonic.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "dihedral_multi_harmonic.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "dihedral_opls.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "dihedral_table.h"
