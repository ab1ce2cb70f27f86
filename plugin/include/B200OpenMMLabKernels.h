/*
 * B200CalcSlicedNonbondedForceKernel -- the drop-in replacement of the reference's platform kernels for
 * SlicedNonbondedForce.  It implements the reference's abstract kernel interface
 * OpenMMLab::CalcSlicedNonbondedForceKernel (openmmapi/include/OpenMMLabKernels.h:29-87) method for method,
 * each a thin forward to one entry point of the C-ABI in include/slicednb.h; everything above this class
 * (SlicedNonbondedForce, SlicedNonbondedForceImpl, the Python wrappers, serialization) is the reference's
 * own, unchanged.
 *
 * Built only where OpenMM (>= 7.7, CI-pinned 8.1) and the reference's headers are installed; in this
 * repository it is compile-checked against the stand-in headers of plugin/mock_openmm (plugin/Makefile).
 */
#ifndef B200_OPENMM_LAB_KERNELS_H_
#define B200_OPENMM_LAB_KERNELS_H_

#include "OpenMMLabKernels.h"
#include "openmm/KernelFactory.h"
#include "slicednb.h"

#include <string>
#include <vector>

namespace OpenMMLab {

class B200CalcSlicedNonbondedForceKernel : public CalcSlicedNonbondedForceKernel {
public:
    B200CalcSlicedNonbondedForceKernel(std::string name, const OpenMM::Platform& platform, int deviceIndex, int flags) :
            CalcSlicedNonbondedForceKernel(name, platform), handle(NULL), deviceIndex(deviceIndex), flags(flags) {
    }
    ~B200CalcSlicedNonbondedForceKernel();
    /* OpenMMLabKernels.h:50 ; reference implementation ReferenceOpenMMLabKernels.cpp:65-191 */
    void initialize(const OpenMM::System& system, const SlicedNonbondedForce& force);
    /* OpenMMLabKernels.h:61 ; ReferenceOpenMMLabKernels.cpp:193-261 */
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal);
    /* OpenMMLabKernels.h:68 ; ReferenceOpenMMLabKernels.cpp:263-312 */
    void copyParametersToContext(OpenMM::ContextImpl& context, const SlicedNonbondedForce& force);
    /* OpenMMLabKernels.h:77,86 ; ReferenceOpenMMLabKernels.cpp:314-330 */
    void getPMEParameters(double& alpha, int& nx, int& ny, int& nz) const;
    void getLJPMEParameters(double& alpha, int& nx, int& ny, int& nz) const;
protected:
    struct Snapshot;                      /* flat copy of the force: the arrays an snb_desc points into */
    void fill(const OpenMM::System& system, const SlicedNonbondedForce& force, Snapshot& s) const;
    snb_handle* handle;
    int deviceIndex, flags;
    std::vector<std::string> globalNames, derivativeNames;
    std::vector<double> positions, forces, globals, derivatives;
};

class B200OpenMMLabKernelFactory : public OpenMM::KernelFactory {
public:
    OpenMM::KernelImpl* createKernelImpl(std::string name, const OpenMM::Platform& platform, OpenMM::ContextImpl& context) const;
};

} // namespace OpenMMLab

/* plugin entry points, as platforms/reference/src/ReferenceOpenMMLabKernelFactory.cpp:20-36 */
extern "C" void registerPlatforms();
extern "C" void registerKernelFactories();
extern "C" void registerOpenMMLabB200KernelFactories();

#endif
