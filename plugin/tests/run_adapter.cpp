/* Link-and-run check of the plugin adapter (plugin/src/B200OpenMMLabKernels.cpp): the reference's OWN API class
 * (openmmapi/src/SlicedNonbondedForce.cpp, compiled unchanged from where it lies) builds the force, the adapter's
 * registerKernelFactories() puts the factory on a Reference platform, the platform creates the kernel by name, and the
 * kernel is driven the way SlicedNonbondedForceImpl drives it (openmmapi/src/SlicedNonbondedForceImpl.cpp:34,132-142,356-367):
 * initialize / execute / copyParametersToContext / getPMEParameters, against closed-form answers.  OpenMM itself is the
 * stand-in of plugin/mock_openmm; the C-ABI behind the adapter is the CPU oracle here (tests/oracle_map.h) and the CUDA
 * library on a GPU box (make run-check-cuda). */
#include "B200OpenMMLabKernels.h"
#include "SlicedNonbondedForce.h"

#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"
#include "snb_constants.h"

#include <cmath>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

using namespace OpenMM;
using namespace OpenMMLab;
using namespace std;

extern "C" void registerOpenMMLabB200CudaKernelFactories() {}      /* the CUDA-platform binding is not part of this check */

static int failures = 0;

static void expect(const char* what, double expected, double found, double tol) {
    const double scale = fabs(expected) > 1.0 ? fabs(expected) : 1.0;
    if (!(fabs(expected-found)/scale <= tol)) {
        printf("FAIL %s: expected %.12g, found %.12g\n", what, expected, found);
        failures++;
    }
}

static void expectThrows(const char* what, bool threw, const string& message, const char* fragment) {
    if (!threw || message.find(fragment) == string::npos) {
        printf("FAIL %s: %s '%s' (wanted '%s')\n", what, threw ? "message" : "no exception", message.c_str(), fragment);
        failures++;
    }
}

struct Harness {
    System system;
    vector<Vec3> positions, velocities, forces;
    Vec3 box[3];
    map<string, double> derivatives;
    ReferencePlatform::PlatformData data;
    ContextImpl* context;
    KernelImpl* impl;
    Harness(int numParticles, Platform& platform) : positions(numParticles), velocities(numParticles), forces(numParticles), impl(NULL) {
        for (int i = 0; i < numParticles; i++)
            system.addParticle(1.0);
        box[0] = Vec3(4, 0, 0); box[1] = Vec3(0, 4, 0); box[2] = Vec3(0, 0, 4);
        system.setDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
        data.positions = &positions; data.velocities = &velocities; data.forces = &forces;
        data.periodicBoxVectors = box; data.energyParameterDerivatives = &derivatives;
        context = new ContextImpl(system, &data);
        impl = platform.createKernelImpl(CalcSlicedNonbondedForceKernel::Name(), *context);     /* Platform::createKernel */
    }
    ~Harness() { delete impl; delete context; }
    CalcSlicedNonbondedForceKernel& kernel() { return dynamic_cast<CalcSlicedNonbondedForceKernel&>(*impl); }
};

static void twoParticles(Platform& platform) {
    SlicedNonbondedForce force(2);
    force.addParticle(1.0, 0.3, 0.5);
    force.addParticle(-1.5, 0.34, 0.8);
    force.setParticleSubset(1, 1);
    force.addGlobalParameter("lambda", 1.0);
    force.addScalingParameter("lambda", 0, 1, true, true);
    force.addScalingParameterDerivative("lambda");
    bool threw = false;
    string message;
    try { force.addScalingParameter("lambda", 1, 0, true, false); }     /* the reference's own clash rule (SlicedNonbondedForce.cpp) */
    catch (const OpenMMException& e) { threw = true; message = e.what(); }
    expectThrows("scaling parameter clash", threw, message, "");
    Harness h(2, platform);
    const double r = 0.41, lambda = 0.37;
    h.positions[0] = Vec3(0.1, 0.2, 0.3);
    h.positions[1] = Vec3(0.1+r, 0.2, 0.3);
    h.forces[0] = Vec3(1, 2, 3);            /* the kernel ADDS its forces */
    h.context->setParameter("lambda", lambda);
    h.kernel().initialize(h.system, force);
    const double energy = h.kernel().execute(*h.context, true, true, true, true);
    const double s = 0.5*(0.3+0.34), e = sqrt(0.5*0.8), s6 = pow(s/r, 6.0), qq = 1.0*-1.5;
    const double ec = SNB_ONE_4PI_EPS0*qq/r, ev = 4.0*e*(s6*s6-s6);
    const double dEdr = -SNB_ONE_4PI_EPS0*qq/(r*r)-4.0*e*(12.0*s6*s6-6.0*s6)/r;
    expect("energy", lambda*(ec+ev), energy, 1e-6);
    expect("dE/dlambda", ec+ev, h.derivatives["lambda"], 1e-6);
    expect("F0x (added to 1)", 1.0+lambda*dEdr, h.forces[0][0], 1e-5);
    expect("F0y (added to 2)", 2.0, h.forces[0][1], 1e-5);
    expect("F1x", -lambda*dEdr, h.forces[1][0], 1e-5);
    /* energy 0 without includeEnergy, derivatives all the same (ReferenceOpenMMLabKernels.cpp:246-260) */
    h.derivatives.clear();
    expect("energy without includeEnergy", 0.0, h.kernel().execute(*h.context, true, false, true, true), 0.0);
    expect("dE/dlambda without includeEnergy", ec+ev, h.derivatives["lambda"], 1e-6);
    /* wrong method for getPMEParameters (ReferenceOpenMMLabKernels.cpp:315-316) */
    threw = false;
    try { double a; int nx, ny, nz; h.kernel().getPMEParameters(a, nx, ny, nz); }
    catch (const OpenMMException& ex) { threw = true; message = ex.what(); }
    expectThrows("getPMEParameters on a NoCutoff force", threw, message, "PME");
    /* copyParametersToContext (ReferenceOpenMMLabKernels.cpp:263-312): new charge, same topology */
    force.setParticleParameters(1, -0.5, 0.34, 0.8);
    h.kernel().copyParametersToContext(*h.context, force);
    const double energy2 = h.kernel().execute(*h.context, false, true, true, true);
    expect("energy after copyParametersToContext", lambda*(SNB_ONE_4PI_EPS0*1.0*-0.5/r+ev), energy2, 1e-6);
    /* ... and a changed particle count is refused with the reference's message (:264-265) */
    force.addParticle(0.0, 1.0, 0.0);
    threw = false;
    try { h.kernel().copyParametersToContext(*h.context, force); }
    catch (const OpenMMException& ex) { threw = true; message = ex.what(); }
    expectThrows("copyParametersToContext with another particle", threw, message, "number of particles");
}

static void fromNonbondedForce(Platform& platform) {
    /* SlicedNonbondedForce(const NonbondedForce&, int) (SlicedNonbondedForce.cpp:34-82): three charges in a periodic
     * reaction-field box, pair 0-1 excluded by an exception */
    NonbondedForce nb;
    nb.setNonbondedMethod(NonbondedForce::CutoffPeriodic);
    nb.setCutoffDistance(1.0);
    nb.setReactionFieldDielectric(50.0);
    nb.setUseDispersionCorrection(false);
    nb.addParticle(1.0, 1.0, 0.0);
    nb.addParticle(-1.0, 1.0, 0.0);
    nb.addParticle(0.5, 1.0, 0.0);
    nb.addException(0, 1, 0.0, 1.0, 0.0);
    SlicedNonbondedForce force(nb, 1);
    if (force.getNumParticles() != 3 || force.getNumExceptions() != 1 || force.getNonbondedMethod() != NonbondedForce::CutoffPeriodic ||
        force.getNonbondedMethodName() != "CutoffPeriodic" || force.getNumSlices() != 1) {
        printf("FAIL copy of a NonbondedForce\n");
        failures++;
    }
    Harness h(3, platform);
    h.positions[0] = Vec3(0.2, 2.0, 2.0);
    h.positions[1] = Vec3(0.7, 2.0, 2.0);       /* 0.5 nm from particle 0: excluded */
    h.positions[2] = Vec3(3.6, 2.0, 2.0);       /* 0.6 nm from particle 0 through the x face, 1.1 nm from particle 1: outside the cutoff */
    h.kernel().initialize(h.system, force);
    const double energy = h.kernel().execute(*h.context, true, true, true, true);
    const double rc = 1.0, r = 0.6, eps = 50.0;
    const double krf = (eps-1.0)/(2.0*eps+1.0)/(rc*rc*rc), crf = 3.0*eps/(2.0*eps+1.0)/rc;
    expect("reaction-field energy through the box face", SNB_ONE_4PI_EPS0*1.0*0.5*(1.0/r+krf*r*r-crf), energy, 1e-6);
    expect("F0x", SNB_ONE_4PI_EPS0*1.0*0.5*(1.0/(r*r)-2.0*krf*r), h.forces[0][0], 1e-5);       /* repelled from the image at x = -0.4 */
    expect("F1 untouched", 0.0, fabs(h.forces[1][0])+fabs(h.forces[1][1])+fabs(h.forces[1][2]), 1e-9);
    /* box smaller than twice the cutoff (ReferenceOpenMMLabKernels.cpp:208-210) */
    h.box[0] = Vec3(1.9, 0, 0);
    bool threw = false;
    string message;
    try { h.kernel().execute(*h.context, true, true, true, true); }
    catch (const OpenMMException& ex) { threw = true; message = ex.what(); }
    expectThrows("box check", threw, message, "twice the nonbonded cutoff");
}

int main() {
    ReferencePlatform platform;
    Platform::registerPlatform(&platform);
    registerKernelFactories();              /* the plugin's own entry point */
    bool threw = false;
    string message;
    {
        System system;
        ReferencePlatform::PlatformData data;
        ContextImpl context(system, &data);
        try { B200OpenMMLabKernelFactory().createKernelImpl("NoSuchKernel", platform, context); }
        catch (const OpenMMException& e) { threw = true; message = e.what(); }
        expectThrows("unknown kernel name", threw, message, "illegal kernel name");
    }
    twoParticles(platform);
    fromNonbondedForce(platform);
    if (failures) {
        printf("%d check(s) failed\n", failures);
        return 1;
    }
    printf("OK\n");
    return 0;
}
