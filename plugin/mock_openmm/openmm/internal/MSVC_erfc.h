// Stand-in for openmm/internal/MSVC_erfc.h: on Linux libm provides erf/erfc.
#include <cmath>
