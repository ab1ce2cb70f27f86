/*
 * See plugin/include/B200OpenMMLabKernels.h.  The kernel keeps host buffers in the layout of the platform it
 * is registered on -- the Reference platform's PlatformData (vector<Vec3> positions / forces, box vectors,
 * energyParameterDerivatives; field use as in platforms/reference/src/ReferenceOpenMMLabKernels.cpp:35-58) --
 * and hands them to snb_execute.  The zero-copy binding to OpenMM's CUDA platform (snb_execute_device_ordered) is
 * B200CudaOpenMMLabKernels.cpp.
 */
#include "B200OpenMMLabKernels.h"

#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"

#include <algorithm>
#include <cstdlib>
#include <map>

using namespace OpenMMLab;
using namespace OpenMM;
using namespace std;

struct B200CalcSlicedNonbondedForceKernel::Snapshot {
    snb_desc desc;
    vector<double> charge, sigma, epsilon, exceptionParams, globalDefaults, particleOffsetScale, exceptionOffsetScale;
    vector<int32_t> subset, exceptionParticles, particleOffsetIndex, exceptionOffsetIndex, scaling, derivativeParams;
};

static void check(int status) {
    if (status != SNB_OK)
        throw OpenMMException(snb_last_error());    /* same messages as the reference throws (:210,265,291,316,325) */
}

B200CalcSlicedNonbondedForceKernel::~B200CalcSlicedNonbondedForceKernel() {
    snb_destroy(handle);
}

void B200CalcSlicedNonbondedForceKernel::fill(const System& system, const SlicedNonbondedForce& force, Snapshot& s) const {
    /* everything ReferenceCalcSlicedNonbondedForceKernel::initialize reads (ReferenceOpenMMLabKernels.cpp:65-191) */
    snb_desc& d = s.desc;
    d = snb_desc();
    d.struct_size = sizeof(snb_desc);
    int n = force.getNumParticles();
    d.num_particles = n;
    d.num_subsets = force.getNumSubsets();
    d.method = (int) force.getNonbondedMethod();           /* same values as CalcSlicedNonbondedForceKernel::NonbondedMethod */
    d.cutoff = force.getCutoffDistance();
    d.use_switching_function = force.getUseSwitchingFunction();
    d.use_dispersion_correction = force.getUseDispersionCorrection();
    d.switching_distance = force.getSwitchingDistance();
    d.reaction_field_dielectric = force.getReactionFieldDielectric();
    d.exceptions_use_pbc = force.getExceptionsUsePeriodicBoundaryConditions();
    d.flags = flags;
    d.ewald_error_tolerance = force.getEwaldErrorTolerance();
    Vec3 box[3];
    system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            d.default_box[3*i+j] = box[i][j];
    int nx, ny, nz;
    force.getPMEParameters(d.ewald_alpha, nx, ny, nz);
    d.pme_grid[0] = nx; d.pme_grid[1] = ny; d.pme_grid[2] = nz;
    force.getLJPMEParameters(d.dispersion_alpha, nx, ny, nz);
    d.dispersion_grid[0] = nx; d.dispersion_grid[1] = ny; d.dispersion_grid[2] = nz;
    d.device_index = deviceIndex;

    s.charge.resize(n); s.sigma.resize(n); s.epsilon.resize(n); s.subset.resize(n);
    for (int i = 0; i < n; i++) {
        force.getParticleParameters(i, s.charge[i], s.sigma[i], s.epsilon[i]);
        s.subset[i] = force.getParticleSubset(i);
    }
    int x = force.getNumExceptions();
    s.exceptionParticles.resize(2*x); s.exceptionParams.resize(3*x);
    for (int i = 0; i < x; i++) {
        int p1, p2;
        force.getExceptionParameters(i, p1, p2, s.exceptionParams[3*i], s.exceptionParams[3*i+1], s.exceptionParams[3*i+2]);
        s.exceptionParticles[2*i] = p1; s.exceptionParticles[2*i+1] = p2;
    }
    int g = force.getNumGlobalParameters();
    map<string, int> index;
    s.globalDefaults.resize(g);
    for (int i = 0; i < g; i++) {
        index[force.getGlobalParameterName(i)] = i;
        s.globalDefaults[i] = force.getGlobalParameterDefaultValue(i);
    }
    int po = force.getNumParticleParameterOffsets();
    s.particleOffsetIndex.resize(2*po); s.particleOffsetScale.resize(3*po);
    for (int i = 0; i < po; i++) {
        string param;
        int particle;
        force.getParticleParameterOffset(i, param, particle, s.particleOffsetScale[3*i], s.particleOffsetScale[3*i+1], s.particleOffsetScale[3*i+2]);
        s.particleOffsetIndex[2*i] = index[param]; s.particleOffsetIndex[2*i+1] = particle;
    }
    int eo = force.getNumExceptionParameterOffsets();
    s.exceptionOffsetIndex.resize(2*eo); s.exceptionOffsetScale.resize(3*eo);
    for (int i = 0; i < eo; i++) {
        string param;
        int exception;
        force.getExceptionParameterOffset(i, param, exception, s.exceptionOffsetScale[3*i], s.exceptionOffsetScale[3*i+1], s.exceptionOffsetScale[3*i+2]);
        s.exceptionOffsetIndex[2*i] = index[param]; s.exceptionOffsetIndex[2*i+1] = exception;
    }
    int k = force.getNumScalingParameters();
    s.scaling.resize(5*k);
    for (int i = 0; i < k; i++) {
        string param;
        int s1, s2;
        bool coulomb, lj;
        force.getScalingParameter(i, param, s1, s2, coulomb, lj);
        s.scaling[5*i] = index[param]; s.scaling[5*i+1] = s1; s.scaling[5*i+2] = s2; s.scaling[5*i+3] = coulomb; s.scaling[5*i+4] = lj;
    }
    int nd = force.getNumScalingParameterDerivatives();
    s.derivativeParams.resize(nd);
    for (int i = 0; i < nd; i++)
        s.derivativeParams[i] = index[force.getScalingParameterDerivativeName(i)];

    d.charge = s.charge.data(); d.sigma = s.sigma.data(); d.epsilon = s.epsilon.data(); d.subset = s.subset.data();
    d.num_exceptions = x; d.exception_particles = s.exceptionParticles.data(); d.exception_params = s.exceptionParams.data();
    d.num_global_parameters = g; d.global_parameter_defaults = s.globalDefaults.data();
    d.num_particle_offsets = po; d.particle_offset_index = s.particleOffsetIndex.data(); d.particle_offset_scale = s.particleOffsetScale.data();
    d.num_exception_offsets = eo; d.exception_offset_index = s.exceptionOffsetIndex.data(); d.exception_offset_scale = s.exceptionOffsetScale.data();
    d.num_scaling_parameters = k; d.scaling_parameters = s.scaling.data();
    d.num_derivatives = nd; d.derivative_parameters = s.derivativeParams.data();
}

void B200CalcSlicedNonbondedForceKernel::initialize(const System& system, const SlicedNonbondedForce& force) {
    Snapshot s;
    fill(system, force, s);
    snb_destroy(handle);
    handle = NULL;
    check(snb_create(&s.desc, &handle));
    globalNames.clear();
    for (int i = 0; i < force.getNumGlobalParameters(); i++)
        globalNames.push_back(force.getGlobalParameterName(i));
    derivativeNames.clear();
    for (int i = 0; i < force.getNumScalingParameterDerivatives(); i++)
        derivativeNames.push_back(force.getScalingParameterDerivativeName(i));
    int n = force.getNumParticles();
    positions.resize(3*n);
    forces.resize(3*n);
    globals.resize(globalNames.size()+1);
    derivatives.resize(derivativeNames.size()+1);
}

double B200CalcSlicedNonbondedForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal) {
    ReferencePlatform::PlatformData* data = reinterpret_cast<ReferencePlatform::PlatformData*>(context.getPlatformData());
    vector<Vec3>& pos = *data->positions;
    vector<Vec3>& force = *data->forces;
    Vec3* boxVectors = data->periodicBoxVectors;
    int n = (int) pos.size();
    for (int i = 0; i < n; i++)
        for (int c = 0; c < 3; c++)
            positions[3*i+c] = pos[i][c];
    double box[9];
    for (int i = 0; i < 3; i++)
        for (int c = 0; c < 3; c++)
            box[3*i+c] = boxVectors[i][c];
    for (size_t i = 0; i < globalNames.size(); i++)
        globals[i] = context.getParameter(globalNames[i]);             /* ReferenceOpenMMLabKernels.cpp:336-340 */
    int flags = (includeForces ? SNB_INCLUDE_FORCES : 0)|(includeEnergy ? SNB_INCLUDE_ENERGY : 0)|
                (includeDirect ? SNB_INCLUDE_DIRECT : 0)|(includeReciprocal ? SNB_INCLUDE_RECIPROCAL : 0);
    std::fill_n(forces.begin(), 3*n, 0.0);
    std::fill(derivatives.begin(), derivatives.end(), 0.0);
    double energy = 0;
    check(snb_execute(handle, positions.data(), box, globals.data(), flags, includeForces ? forces.data() : NULL, &energy, NULL, derivatives.data()));
    if (includeForces)
        for (int i = 0; i < n; i++)
            force[i] += Vec3(forces[3*i], forces[3*i+1], forces[3*i+2]);
    map<string, double>& energyParamDerivs = *data->energyParameterDerivatives;      /* :252-258 */
    for (size_t i = 0; i < derivativeNames.size(); i++)
        energyParamDerivs[derivativeNames[i]] += derivatives[i];
    return energy;
}

void B200CalcSlicedNonbondedForceKernel::copyParametersToContext(ContextImpl& context, const SlicedNonbondedForce& force) {
    Snapshot s;
    fill(context.getSystem(), force, s);
    check(snb_update_parameters(handle, &s.desc));
}

void B200CalcSlicedNonbondedForceKernel::getPMEParameters(double& alpha, int& nx, int& ny, int& nz) const {
    check(snb_get_pme_parameters(handle, &alpha, &nx, &ny, &nz));
}

void B200CalcSlicedNonbondedForceKernel::getLJPMEParameters(double& alpha, int& nx, int& ny, int& nz) const {
    check(snb_get_ljpme_parameters(handle, &alpha, &nx, &ny, &nz));
}

/* ---- plugin registration: platforms/reference/src/ReferenceOpenMMLabKernelFactory.cpp:20-45 ---- */

extern "C" void registerPlatforms() {
}

extern "C" void registerKernelFactories() {
    /* Registering a factory under a kernel name replaces the platform's earlier one, so loading this
     * plugin AFTER the reference's takes over "CalcSlicedNonbondedForce" on the Reference platform and
     * leaves the reference's other kernels (ExtendedCustomCVForce) where they are. */
    for (int i = 0; i < Platform::getNumPlatforms(); i++) {
        Platform& platform = Platform::getPlatform(i);
        if (dynamic_cast<ReferencePlatform*>(&platform) != NULL)
            platform.registerKernelFactory(CalcSlicedNonbondedForceKernel::Name(), new B200OpenMMLabKernelFactory());
    }
}

extern "C" void registerOpenMMLabB200CudaKernelFactories();     /* B200CudaOpenMMLabKernels.cpp: the zero-copy binding */

extern "C" void registerOpenMMLabB200KernelFactories() {
    registerKernelFactories();
    registerOpenMMLabB200CudaKernelFactories();
}

KernelImpl* B200OpenMMLabKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name == CalcSlicedNonbondedForceKernel::Name()) {
        /* the platform properties of the reference's CUDA platform (CudaOpenMMLabKernels.cpp:483,496), read
         * from the environment here because the Reference platform defines no properties of its own */
        const char* device = getenv("SNB_DEVICE_INDEX");
        const char* precision = getenv("SNB_PRECISION");
        const char* deterministic = getenv("SNB_DETERMINISTIC_FORCES");
        int flags = 0;
        if (precision != NULL && string(precision) == "double")
            flags |= SNB_FLAG_DOUBLE_PRECISION;
        if (deterministic != NULL && string(deterministic) != "0" && string(deterministic) != "false")
            flags |= SNB_FLAG_DETERMINISTIC;
        const char* directKernel = getenv("SNB_DIRECT_KERNEL");     /* auto (by system size) | packed | general */
        if (directKernel != NULL && string(directKernel) == "packed")
            flags |= SNB_FLAG_DIRECT_PACKED;
        else if (directKernel != NULL && string(directKernel) == "general")
            flags |= SNB_FLAG_DIRECT_GENERAL;
        return new B200CalcSlicedNonbondedForceKernel(name, platform, device != NULL ? atoi(device) : 0, flags);
    }
    throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '")+name+"'").c_str());
}
