/* see Rinternals.h in this directory -- test stand-in for R's headers */
#include <math.h>
#include <stdlib.h>
#include <string.h>
