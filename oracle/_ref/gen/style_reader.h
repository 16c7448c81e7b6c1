#include "reader_native.h"
#include "reader_xyz.h"
