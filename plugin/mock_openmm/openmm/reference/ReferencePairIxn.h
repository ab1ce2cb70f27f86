// Stand-in for openmm/reference/ReferencePairIxn.h (included, nothing used on this path).
#include "openmm/Vec3.h"
#include <vector>
