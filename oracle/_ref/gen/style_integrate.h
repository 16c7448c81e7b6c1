#include "respa.h"
#include "verlet.h"
