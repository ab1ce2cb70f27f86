/* Differential check of the oracle's RESTATED kernel glue (oracle/harness.cpp: initialize / execute / computeParameters /
 * copyParametersToContext of ReferenceCalcSlicedNonbondedForceKernel, which cannot compile without OpenMM) against the REAL glue:
 * the reference's own platforms/reference/src/ReferenceOpenMMLabKernels.cpp, compiled unchanged on the stand-in OpenMM of
 * plugin/mock_openmm.  Two Reference platforms in one program -- one with the reference's own kernel, one with the plugin adapter
 * and the oracle behind the C-ABI -- evaluate the same random systems (three subsets, exceptions from bonds, scaling parameters
 * with derivatives, particle and exception parameter offsets, switching function, dispersion correction) with all six
 * nonbonded methods, before and after parameter changes; energies, every force component and every derivative must agree to
 * 1e-9 of their scale (the arithmetic underneath is the same code, only the order of a few sums differs). */
#include "B200OpenMMLabKernels.h"
#include "ReferenceOpenMMLabKernels.h"
#include "SlicedNonbondedForce.h"

#include "openmm/Context.h"
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
#include "openmm/VerletIntegrator.h"
#include "openmm/reference/ReferencePlatform.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

using namespace OpenMM;
using namespace OpenMMLab;
using namespace std;

extern "C" void registerOpenMMLabB200CudaKernelFactories() {}

const std::string& OpenMMLab::ExtendedCustomCVForce::getGlobalParameterName(int) const { throw OpenMMException("not part of this program"); }
const std::string& OpenMMLab::ExtendedCustomCVForce::getEnergyParameterDerivativeName(int) const { throw OpenMMException("not part of this program"); }

class OwnFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl&) const {
        return new ReferenceCalcSlicedNonbondedForceKernel(name, platform);
    }
};

static int failures = 0, comparisons = 0;
static double worst = 0;
static double tolerance = 1e-9;     /* DIFF_TOL overrides (the CUDA library behind the adapter: 1e-7 in double precision, 1e-4 mixed) */

static void compare(const char* what, const string& label, double a, double b, double scale) {
    comparisons++;
    const double err = fabs(a-b)/(scale > 1.0 ? scale : 1.0);
    if (err > worst)
        worst = err;
    if (!(err <= tolerance)) {
        if (failures < 20)
            printf("FAIL %s %s: own %.15g, oracle %.15g\n", label.c_str(), what, a, b);
        failures++;
    }
}

static void compareStates(const string& label, Context& own, Context& oracle, int groups) {
    const int types = State::Forces|State::Energy|State::ParameterDerivatives;
    State a = own.getState(types, false, groups), b = oracle.getState(types, false, groups);
    compare("energy", label, a.getPotentialEnergy(), b.getPotentialEnergy(), fabs(a.getPotentialEnergy()));
    double fscale = 0;
    for (size_t i = 0; i < a.getForces().size(); i++)
        fscale = max(fscale, sqrt(a.getForces()[i].dot(a.getForces()[i])));
    for (size_t i = 0; i < a.getForces().size(); i++)
        for (int c = 0; c < 3; c++)
            compare("force", label, a.getForces()[i][c], b.getForces()[i][c], fscale);
    if (a.getEnergyParameterDerivatives().size() != b.getEnergyParameterDerivatives().size()) {
        printf("FAIL %s: different sets of derivatives\n", label.c_str());
        failures++;
    }
    for (map<string, double>::const_iterator d = a.getEnergyParameterDerivatives().begin(); d != a.getEnergyParameterDerivatives().end(); ++d)
        compare(("derivative "+d->first).c_str(), label, d->second, b.getEnergyParameterDerivatives().at(d->first), fabs(d->second));
}

static SlicedNonbondedForce* buildForce(mt19937& rng, int n, NonbondedForce::NonbondedMethod method, double boxSize, bool useSwitch, vector<pair<int, int> >& bonds) {
    uniform_real_distribution<double> u(0.0, 1.0);
    SlicedNonbondedForce* force = new SlicedNonbondedForce(3);
    force->setNonbondedMethod((SlicedNonbondedForce::NonbondedMethod) method);
    force->setCutoffDistance(1.0);
    force->setReactionFieldDielectric(60.0);
    force->setUseDispersionCorrection(true);
    if (useSwitch) {
        force->setUseSwitchingFunction(true);
        force->setSwitchingDistance(0.8);
    }
    /* explicit, cubic PME grids (the reference strides its subset grids by nx where it means nz: ReferencePME.cpp:682) */
    force->setPMEParameters(3.2, 20, 20, 20);
    force->setLJPMEParameters(2.9, 16, 16, 16);
    force->setEwaldErrorTolerance(1e-4);
    double total = 0;
    for (int i = 0; i < n; i++) {
        double q = i == n-1 ? -total : u(rng)-0.5;
        total += q;
        force->addParticle(q, 0.2+0.15*u(rng), 0.1+u(rng));
        force->setParticleSubset(i, i < n/10 ? 1 : (i < n/4 ? 2 : 0));
    }
    for (int i = 0; i+1 < n/5; i++)
        bonds.push_back(make_pair(i, i+1));
    force->createExceptionsFromBonds(bonds, 0.8333, 0.5);
    force->addGlobalParameter("lambdaC", 0.7);
    force->addGlobalParameter("lambdaV", 0.4);
    force->addGlobalParameter("lambdaB", 0.9);
    force->addGlobalParameter("offset", 0.3);
    force->addScalingParameter("lambdaC", 0, 1, true, false);
    force->addScalingParameter("lambdaV", 0, 1, false, true);
    force->addScalingParameter("lambdaB", 1, 2, true, true);
    force->addScalingParameterDerivative("lambdaC");
    force->addScalingParameterDerivative("lambdaB");
    force->addParticleParameterOffset("offset", 3, 0.2, 0.05, 0.3);
    force->addParticleParameterOffset("offset", n/2, -0.2, 0.0, 0.1);
    for (int e = 0; e < force->getNumExceptions(); e++) {
        int p1, p2;
        double qq, sigma, epsilon;
        force->getExceptionParameters(e, p1, p2, qq, sigma, epsilon);
        if (epsilon != 0.0) {               /* a 1-4: give it an offset */
            force->addExceptionParameterOffset("offset", e, 0.1, 0.02, 0.2);
            break;
        }
    }
    return force;
}

int main() {
    if (getenv("DIFF_TOL") != NULL)
        tolerance = atof(getenv("DIFF_TOL"));
    ReferencePlatform own, oracle;
    own.registerKernelFactory(CalcSlicedNonbondedForceKernel::Name(), new OwnFactory());
    oracle.registerKernelFactory(CalcSlicedNonbondedForceKernel::Name(), new B200OpenMMLabKernelFactory());
    const NonbondedForce::NonbondedMethod methods[6] = {NonbondedForce::NoCutoff, NonbondedForce::CutoffNonPeriodic, NonbondedForce::CutoffPeriodic,
                                                        NonbondedForce::Ewald, NonbondedForce::PME, NonbondedForce::LJPME};
    const char* names[6] = {"NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "Ewald", "PME", "LJPME"};
    try {
        for (int m = 0; m < 6; m++)
            for (int variant = 0; variant < 3; variant++) {       /* 0 plain, 1 switching function + reciprocal-space group, 2 triclinic box */
                if (variant == 2 && methods[m] == NonbondedForce::Ewald)
                    continue;           /* (the reference's Impl refuses Ewald in a non-rectangular box) */
                mt19937 rng(1234+10*m+variant);
                uniform_real_distribution<double> u(0.0, 1.0);
                const int n = 240;
                const double boxSize = 3.0;
                const bool useSwitch = variant == 1 && m != 0;
                /* two systems with identical forces (a System owns its forces) */
                System systemA, systemB;
                vector<pair<int, int> > bondsA, bondsB;
                mt19937 rngA = rng, rngB = rng;
                SlicedNonbondedForce* forceA = buildForce(rngA, n, methods[m], boxSize, useSwitch, bondsA);
                SlicedNonbondedForce* forceB = buildForce(rngB, n, methods[m], boxSize, useSwitch, bondsB);
                if (variant == 1) {         /* reciprocal space in a force group of its own */
                    forceA->setReciprocalSpaceForceGroup(3);
                    forceB->setReciprocalSpaceForceGroup(3);
                }
                System* systems[2] = {&systemA, &systemB};
                for (int s = 0; s < 2; s++) {
                    for (int i = 0; i < n; i++)
                        systems[s]->addParticle(1.0);
                    if (variant == 2)
                        systems[s]->setDefaultPeriodicBoxVectors(Vec3(boxSize, 0, 0), Vec3(0.4, boxSize, 0), Vec3(-0.3, 0.5, boxSize));
                    else
                        systems[s]->setDefaultPeriodicBoxVectors(Vec3(boxSize, 0, 0), Vec3(0, boxSize, 0), Vec3(0, 0, boxSize));
                }
                systemA.addForce(forceA);
                systemB.addForce(forceB);
                VerletIntegrator integratorA(0.001), integratorB(0.001);
                Context contextA(systemA, integratorA, own), contextB(systemB, integratorB, oracle);
                /* bonded neighbours 0.15 nm apart along a random walk, everything else on a jittered lattice */
                vector<Vec3> positions(n);
                const int side = 7;
                for (int i = 0; i < n; i++)
                    positions[i] = Vec3(((i%side)+0.2+0.6*u(rng))*boxSize/side, (((i/side)%side)+0.2+0.6*u(rng))*boxSize/side,
                                        ((i/(side*side))+0.2+0.6*u(rng))*boxSize/side);
                contextA.setPositions(positions);
                contextB.setPositions(positions);
                const string label = string(names[m])+(variant == 1 ? "/switch+recipGroup" : variant == 2 ? "/triclinic" : "");
                compareStates(label, contextA, contextB, 0xFFFFFFFF);
                compareStates(label+" [group 0 only]", contextA, contextB, 1<<0);
                compareStates(label+" [group 3 only]", contextA, contextB, 1<<3);
                /* new values of the scaling and offset parameters */
                const char* params[4] = {"lambdaC", "lambdaV", "lambdaB", "offset"};
                for (int p = 0; p < 4; p++) {
                    const double v = u(rng);
                    contextA.setParameter(params[p], v);
                    contextB.setParameter(params[p], v);
                }
                compareStates(label+" [new parameters]", contextA, contextB, 0xFFFFFFFF);
                /* updateParametersInContext: new particle and exception parameters, same topology */
                SlicedNonbondedForce* forces[2] = {forceA, forceB};
                Context* contexts[2] = {&contextA, &contextB};
                for (int s = 0; s < 2; s++) {
                    mt19937 r2(99+m);
                    uniform_real_distribution<double> w(0.0, 1.0);
                    for (int i = 0; i < n; i += 7) {
                        double q, sigma, epsilon;
                        forces[s]->getParticleParameters(i, q, sigma, epsilon);
                        forces[s]->setParticleParameters(i, q*(0.5+w(r2)), sigma*(0.9+0.2*w(r2)), epsilon*(0.5+w(r2)));
                    }
                    for (int e = 0; e < forces[s]->getNumExceptions(); e += 3) {
                        int p1, p2;
                        double qq, sigma, epsilon;
                        forces[s]->getExceptionParameters(e, p1, p2, qq, sigma, epsilon);
                        if (epsilon != 0.0)
                            forces[s]->setExceptionParameters(e, p1, p2, qq*(0.5+w(r2)), sigma, epsilon*(0.5+w(r2)));
                    }
                    forces[s]->updateParametersInContext(*contexts[s]);
                }
                compareStates(label+" [updateParametersInContext]", contextA, contextB, 0xFFFFFFFF);
            }
    }
    catch (const exception& e) {
        printf("exception: %s\n", e.what());
        return 1;
    }
    printf("%d comparisons, worst relative difference %.3g\n", comparisons, worst);
    if (failures) {
        printf("%d FAILED\n", failures);
        return 1;
    }
    printf("OK\n");
    return 0;
}
