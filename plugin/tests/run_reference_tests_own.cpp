/* The reference's OWN test program on the reference's OWN Reference-platform kernel (platforms/reference/src/
 * ReferenceOpenMMLabKernels.cpp + ReferenceOpenMMLabKernelFactory.cpp + ReferenceSlicedLJCoulombIxn.cpp + ReferenceSlicedLJCoulomb14.cpp +
 * ReferencePME.cpp, all compiled unchanged) on the stand-in OpenMM: no code of this repository on the path but the stand-ins.
 * It shows that the stand-in OpenMM carries the reference faithfully (the same program then runs on the adapter + oracle /
 * CUDA library), and it is the yardstick tests/test_abi.py compares the oracle's restated kernel glue with. */
#include "ExtendedCustomCVForce.h"
#include "openmm/OpenMMException.h"

/* the ExtendedCustomCVForce kernel in the reference's kernel file is never created here; the two accessors it names need a body to link */
const std::string& OpenMMLab::ExtendedCustomCVForce::getGlobalParameterName(int) const { throw OpenMM::OpenMMException("not part of this program"); }
const std::string& OpenMMLab::ExtendedCustomCVForce::getEnergyParameterDerivativeName(int) const { throw OpenMM::OpenMMException("not part of this program"); }

#include "ReferenceOpenMMLabTests.h"
#include "TestSlicedNonbondedForce.h"

void runPlatformTests() {
}
