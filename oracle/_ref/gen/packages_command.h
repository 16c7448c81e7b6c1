#undef PACKAGE
#define PACKAGE "PLUGIN"
#include "plugin.h"
