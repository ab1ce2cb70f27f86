/* The reference's OWN test program for SlicedNonbondedForce -- tests/TestSlicedNonbondedForce.h with its main(), compiled
 * UNCHANGED from where it lies, exactly as platforms/reference/tests/TestReferenceSlicedNonbondedForce.cpp builds it -- run
 * through the plugin adapter on the stand-in OpenMM of plugin/mock_openmm (plugin/Makefile, target run-reference-tests).
 * The reference's harness header expects registerOpenMMLabReferenceKernelFactories(); here that registers THIS plugin. */
#include "B200OpenMMLabKernels.h"

extern "C" void registerOpenMMLabB200CudaKernelFactories() {}      /* the CUDA-platform binding is not part of this program */
extern "C" void registerOpenMMLabReferenceKernelFactories() { registerKernelFactories(); }

#include "ReferenceOpenMMLabTests.h"
#include "TestSlicedNonbondedForce.h"

void runPlatformTests() {
}
