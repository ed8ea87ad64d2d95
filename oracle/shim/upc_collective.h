/* upc_collective.h -- see upc.h in this directory (test oracle shim). */
#include "upc.h"
