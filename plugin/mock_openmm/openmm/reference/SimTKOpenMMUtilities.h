// Stand-in for openmm/reference/SimTKOpenMMUtilities.h (included, nothing used on this path).
#include "openmm/reference/SimTKOpenMMRealType.h"
