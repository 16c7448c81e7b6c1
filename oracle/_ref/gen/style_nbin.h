#include "nbin_multi.h"
#include "nbin_standard.h"
