#undef PACKAGE
#define PACKAGE "OBMD"
#include "atom_vec_adress.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "atom_vec_angle.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "atom_vec_bond.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "atom_vec_full.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "atom_vec_molecular.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "atom_vec_template.h"
