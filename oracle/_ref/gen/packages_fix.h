#undef PACKAGE
#define PACKAGE "OBMD"
#include "fix_obmd_merged.h"
#undef PACKAGE
#define PACKAGE "MOLECULE"
#include "fix_cmap.h"
