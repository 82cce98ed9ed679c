/* rosette.c -- oracle (TEST INFRASTRUCTURE, see oracle.h): strain rosettes (placeholder). */
#include "oracle.h"
