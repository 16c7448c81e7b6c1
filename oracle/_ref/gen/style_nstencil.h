#include "nstencil_bin.h"
#include "nstencil_ghost_bin.h"
#include "nstencil_multi.h"
#include "nstencil_multi_old.h"
