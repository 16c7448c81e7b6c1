#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "angle_charmm.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "angle_cosine.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "angle_cosine_squared.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "angle_harmonic.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "angle_table.h"
