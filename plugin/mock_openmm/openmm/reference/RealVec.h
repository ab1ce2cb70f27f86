/* STAND-IN for openmm/reference/RealVec.h. */
#ifndef MOCK_OPENMM_REALVEC_H_
#define MOCK_OPENMM_REALVEC_H_
#include "../Vec3.h"
namespace OpenMM { typedef Vec3 RealVec; }
#endif
