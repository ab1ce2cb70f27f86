/*
 * See plugin/include/B200CudaOpenMMLabKernels.h.  Compile-checked against the reference's own headers and stand-in OpenMM
 * CUDA-platform headers (plugin/mock_openmm/openmm/cuda); the device entry point it drives is tested on the GPU by
 * tests/test_device_ordered.py with the same layout (reordered atoms, padded buffers, a force buffer that is added to).
 */
#include "B200CudaOpenMMLabKernels.h"

#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

#include <cstdlib>
#include <map>

using namespace OpenMMLab;
using namespace OpenMM;
using namespace std;

static void check(int status) {
    if (status != SNB_OK)
        throw OpenMMException(snb_last_error());
}

/* Which atoms / exceptions OpenMM may treat as interchangeable when it reorders atoms by molecule: the criteria of the
 * reference's CUDA kernel (CudaOpenMMLabKernels.cpp:35-73) -- equal parameters AND equal subset; exceptions equal in
 * parameters and slice. */
class B200CudaCalcSlicedNonbondedForceKernel::ForceInfo : public CudaForceInfo {
public:
    ForceInfo(const SlicedNonbondedForce& force) : force(force) {
    }
    bool areParticlesIdentical(int particle1, int particle2) {
        double q1, s1, e1, q2, s2, e2;
        force.getParticleParameters(particle1, q1, s1, e1);
        force.getParticleParameters(particle2, q2, s2, e2);
        return q1 == q2 && s1 == s2 && e1 == e2 && force.getParticleSubset(particle1) == force.getParticleSubset(particle2);
    }
    int getNumParticleGroups() {
        return force.getNumExceptions();
    }
    void getParticlesInGroup(int index, vector<int>& particles) {
        int p1, p2;
        double qq, s, e;
        force.getExceptionParameters(index, p1, p2, qq, s, e);
        particles.resize(2);
        particles[0] = p1;
        particles[1] = p2;
    }
    bool areGroupsIdentical(int group1, int group2) {
        int a1, a2, b1, b2;
        double qa, sa, ea, qb, sb, eb;
        force.getExceptionParameters(group1, a1, a2, qa, sa, ea);
        force.getExceptionParameters(group2, b1, b2, qb, sb, eb);
        return qa == qb && sa == sb && ea == eb && slice(a1, a2) == slice(b1, b2);
    }
private:
    int slice(int p1, int p2) const {
        int i = force.getParticleSubset(p1), j = force.getParticleSubset(p2);
        return i > j ? i*(i+1)/2+j : j*(j+1)/2+i;
    }
    const SlicedNonbondedForce& force;
};

void B200CudaCalcSlicedNonbondedForceKernel::initialize(const System& system, const SlicedNonbondedForce& force) {
    cu.setAsCurrent();
    B200CalcSlicedNonbondedForceKernel::initialize(system, force);      /* snapshot -> snb_create on cu's device */
    if (listSkin > 0)
        check(snb_set_list_skin(handle, listSkin));
    cu.addForce(new ForceInfo(force));
}

double B200CudaCalcSlicedNonbondedForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal) {
    cu.setAsCurrent();
    /* the platform's positions must be complete before the core's own stream reads them */
    if (cuStreamSynchronize(cu.getCurrentStream()) != CUDA_SUCCESS)
        throw OpenMMException("B200CudaCalcSlicedNonbondedForceKernel: the CUDA platform's stream failed");
    Vec3 a, b, c;
    cu.getPeriodicBoxVectors(a, b, c);
    double box[9] = {a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]};
    for (size_t i = 0; i < globalNames.size(); i++)
        globals[i] = context.getParameter(globalNames[i]);
    int flags = (includeForces ? SNB_INCLUDE_FORCES : 0)|(includeEnergy ? SNB_INCLUDE_ENERGY : 0)|
                (includeDirect ? SNB_INCLUDE_DIRECT : 0)|(includeReciprocal ? SNB_INCLUDE_RECIPROCAL : 0);
    snb_device_layout layout = snb_device_layout();
    layout.struct_size = sizeof(snb_device_layout);
    layout.padded_num_atoms = cu.getPaddedNumAtoms();
    layout.positions_double = cu.getUseDoublePrecision();
    layout.posq = (const void*) cu.getPosq().getDevicePointer();
    layout.posq_correction = cu.getUseMixedPrecision() ? (const void*) cu.getPosqCorrection().getDevicePointer() : NULL;
    layout.atom_index = (const int32_t*) cu.getAtomIndexArray().getDevicePointer();
    layout.forces = includeForces ? (void*) cu.getForce().getDevicePointer() : NULL;
    /* returns after the core's stream has finished: the platform's next kernel sees the forces */
    check(snb_execute_device_ordered(handle, &layout, box, globals.data(), flags));
    double energy = 0;
    std::fill(derivatives.begin(), derivatives.end(), 0.0);
    check(snb_read_energies(handle, flags, &energy, NULL, derivatives.data()));
    map<string, double>& energyParamDerivs = cu.getEnergyParamDerivWorkspace();
    for (size_t i = 0; i < derivativeNames.size(); i++)
        energyParamDerivs[derivativeNames[i]] += derivatives[i];
    return energy;
}

KernelImpl* B200CudaOpenMMLabKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name == CalcSlicedNonbondedForceKernel::Name()) {
        CudaPlatform::PlatformData& data = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData());
        if (data.contexts.size() != 1)
            throw OpenMMException("B200 SlicedNonbondedForce kernel: one device per OpenMM context (multi-GPU runs use one process per GPU, snb_comm_init)");
        CudaContext& cu = *data.contexts[0];
        /* arithmetic follows the platform's Precision property: single / mixed -> the core's mixed mode, double -> double */
        int flags = cu.getUseDoublePrecision() ? SNB_FLAG_DOUBLE_PRECISION : 0;
        const char* deterministic = getenv("SNB_DETERMINISTIC_FORCES");
        if (deterministic != NULL && string(deterministic) != "0" && string(deterministic) != "false")
            flags |= SNB_FLAG_DETERMINISTIC;
        /* OpenMM's CUDA platform pads its neighbour list by 10 % of the cutoff and rebuilds when an atom has moved half of
         * that; SNB_LIST_SKIN (nm) asks the core for the same behaviour (0 / unset: rebuilt every evaluation) */
        const char* skin = getenv("SNB_LIST_SKIN");
        return new B200CudaCalcSlicedNonbondedForceKernel(name, platform, cu, flags, skin != NULL ? atof(skin) : 0.0);
    }
    throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '")+name+"'").c_str());
}

/* Loaded AFTER the reference's own CUDA plugin: takes over "CalcSlicedNonbondedForce" on the CUDA platform and leaves the
 * reference's other kernels where they are (platforms/cuda/src/CudaOpenMMLabKernelFactory.cpp:19-57). */
extern "C" void registerOpenMMLabB200CudaKernelFactories() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        platform.registerKernelFactory(CalcSlicedNonbondedForceKernel::Name(), new B200CudaOpenMMLabKernelFactory());
    }
    catch (const std::exception&) {
        /* no CUDA platform in this OpenMM: nothing to register */
    }
}
