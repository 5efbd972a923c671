// stand-in for <gsl/gsl_math.h> (GSL not installed); the real header pulls in <math.h>
#include <math.h>
#include <cmath>
