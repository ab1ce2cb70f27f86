/* STAND-IN for openmm/reference/ReferenceBondForce.h (included by the reference's kernel file, nothing of it used on this path). */
#ifndef MOCK_OPENMM_REFERENCEBONDFORCE_H_
#define MOCK_OPENMM_REFERENCEBONDFORCE_H_
#include "../Vec3.h"
#endif
