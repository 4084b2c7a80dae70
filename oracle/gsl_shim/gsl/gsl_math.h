/* Minimal GSL-API shim header (test infrastructure; see ../gsl_shim_all.h). */
#include "../gsl_shim_all.h"
