#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "improper_cvff.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "improper_harmonic.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "improper_umbrella.h"
