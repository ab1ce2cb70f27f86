/*
 * B200CudaCalcSlicedNonbondedForceKernel -- the same kernel registered on OpenMM's CUDA platform, ZERO-COPY: positions are
 * read from, and forces added into, the platform's own device buffers.  It replaces
 * CudaCalcSlicedNonbondedForceKernel (platforms/cuda/include/CudaOpenMMLabKernels.h, platforms/cuda/src/
 * CudaOpenMMLabKernels.cpp:35-73 ForceInfo, :736-758 hand-over of posq / parameters, :1114-1203 parameter updates) and
 * does not use CudaNonbondedUtilities at all: the B200 core owns its sort, tile list and PME.
 *
 *   positions   cu.getPosq() [+ cu.getPosqCorrection() in mixed precision], float4 / double4, in the platform's reordered
 *               atom order, padded to cu.getPaddedNumAtoms()
 *   atom order  cu.getAtomIndexArray(): original index of the atom in every slot (it changes when OpenMM reorders)
 *   forces      cu.getForce(): long long [3][paddedNumAtoms], 2^32 fixed point (platforms/common/src/kernels/pme.cc:399-405),
 *               ADDED to by snb_execute_device_ordered
 *
 * One CudaContext (single device per OpenMM context).  OpenMM's multi-device mode (CudaParallelOpenMMLabKernels.cpp:19-53:
 * several contexts in one process) is replaced by the core's own one-process-per-GPU partition (snb_comm_init).
 */
#ifndef B200_CUDA_OPENMM_LAB_KERNELS_H_
#define B200_CUDA_OPENMM_LAB_KERNELS_H_

#include "B200OpenMMLabKernels.h"
#include "openmm/cuda/CudaContext.h"

namespace OpenMMLab {

class B200CudaCalcSlicedNonbondedForceKernel : public B200CalcSlicedNonbondedForceKernel {
public:
    B200CudaCalcSlicedNonbondedForceKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu, int flags, double listSkin) :
            B200CalcSlicedNonbondedForceKernel(name, platform, cu.getDeviceIndex(), flags), cu(cu), listSkin(listSkin) {
    }
    void initialize(const OpenMM::System& system, const SlicedNonbondedForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal);
private:
    class ForceInfo;
    OpenMM::CudaContext& cu;
    double listSkin;
};

class B200CudaOpenMMLabKernelFactory : public OpenMM::KernelFactory {
public:
    OpenMM::KernelImpl* createKernelImpl(std::string name, const OpenMM::Platform& platform, OpenMM::ContextImpl& context) const;
};

} // namespace OpenMMLab

extern "C" void registerOpenMMLabB200CudaKernelFactories();

#endif
