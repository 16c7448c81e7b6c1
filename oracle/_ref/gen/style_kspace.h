#include "kspace_deprecated.h"
