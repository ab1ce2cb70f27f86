// Stand-in for openmm/reference/SimTKOpenMMRealType.h: the three symbols the
// reference path uses (ReferenceSlicedLJCoulombIxn.cpp:184-186,394; ReferencePME.cpp:425).
#ifndef SNB_SHIM_REALTYPE_H_
#define SNB_SHIM_REALTYPE_H_
#include <cmath>
#include "snb_constants.h"
#define ONE_4PI_EPS0 SNB_ONE_4PI_EPS0
#define PI_M 3.14159265358979323846
#define EXP exp
#define SQRT sqrt
#endif
