/* STAND-IN for openmm/reference/SimTKOpenMMRealType.h: the constants the reference's tests use.  ONE_4PI_EPS0 is the value
 * oracle and product share (include/snb_constants.h). */
#ifndef MOCK_OPENMM_REALTYPE_H_
#define MOCK_OPENMM_REALTYPE_H_
#include "snb_constants.h"
#include <cmath>
#define ONE_4PI_EPS0 SNB_ONE_4PI_EPS0
#define SQRT_TWO 1.41421356237309504
#define PI_M 3.141592653589793238462643383279502884197169399375
#define SQRT std::sqrt
#define EXP std::exp
#define ERFC erfc
#endif
