/* STAND-IN for openmm/internal/NonbondedForceImpl.h: the base the reference's SlicedNonbondedForceImpl derives from (it
 * overrides everything but the default parameters: NonbondedForceImpl::getDefaultParameters = the force's global parameters). */
#ifndef MOCK_OPENMM_NONBONDEDFORCEIMPL_H_
#define MOCK_OPENMM_NONBONDEDFORCEIMPL_H_
#include "ForceImpl.h"
#include "../NonbondedForce.h"
#include "../System.h"
#include "ContextImpl.h"
#include <cmath>
namespace OpenMM {
class NonbondedForceImpl : public ForceImpl {
public:
    NonbondedForceImpl(const NonbondedForce& owner) : owner(owner) {}
    void initialize(ContextImpl&) { throw OpenMMException("mock OpenMM: NonbondedForceImpl::initialize"); }
    const NonbondedForce& getOwner() const { return owner; }
    double calcForcesAndEnergy(ContextImpl&, bool, bool, int) { throw OpenMMException("mock OpenMM: NonbondedForceImpl::calcForcesAndEnergy"); }
    std::map<std::string, double> getDefaultParameters() {
        std::map<std::string, double> parameters;
        for (int i = 0; i < owner.getNumGlobalParameters(); i++)
            parameters[owner.getGlobalParameterName(i)] = owner.getGlobalParameterDefaultValue(i);
        return parameters;
    }
    std::vector<std::string> getKernelNames() { return std::vector<std::string>(); }
    /* OpenMM 8.1's rules (restated from their published form; explicit parameters win):
     * alpha = sqrt(-ln(2 tol))/rc; PME grid n = max(ceil(2 alpha L/(3 tol^(1/5))), 6), the dispersion grid without the factor 2;
     * Ewald kmax = the smallest odd k with 0.05 sqrt(L alpha) k exp(-(pi k/(L alpha))^2) < tol */
    static void calcPMEParameters(const System& system, const NonbondedForce& force, double& alpha, int& xsize, int& ysize, int& zsize, bool lj) {
        if (lj) force.getLJPMEParameters(alpha, xsize, ysize, zsize);
        else force.getPMEParameters(alpha, xsize, ysize, zsize);
        if (alpha == 0.0) {
            Vec3 box[3];
            system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
            const double tol = force.getEwaldErrorTolerance();
            alpha = (1.0/force.getCutoffDistance())*std::sqrt(-std::log(2.0*tol));
            int* size[3] = {&xsize, &ysize, &zsize};
            for (int k = 0; k < 3; k++) {
                const double L = box[k][k];
                const int n = lj ? (int) std::ceil(alpha*L/(3*std::pow(tol, 0.2))) : (int) std::ceil(2*alpha*L/(3*std::pow(tol, 0.2)));
                *size[k] = n > 6 ? n : 6;
            }
        }
    }
    static void calcEwaldParameters(const System& system, const NonbondedForce& force, double& alpha, int& kmaxx, int& kmaxy, int& kmaxz) {
        Vec3 box[3];
        system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
        const double tol = force.getEwaldErrorTolerance();
        alpha = (1.0/force.getCutoffDistance())*std::sqrt(-std::log(2.0*tol));
        int* kmax[3] = {&kmaxx, &kmaxy, &kmaxz};
        for (int k = 0; k < 3; k++) {
            const double width = box[k][k];
            int arg = 10;
            double value = ewaldError(width, alpha, tol, arg);
            if (value > 0.0) {
                while (value > 0.0 && arg > 0)
                    value = ewaldError(width, alpha, tol, --arg);
                arg++;
            }
            else
                while (value < 0.0)
                    value = ewaldError(width, alpha, tol, ++arg);
            *kmax[k] = arg%2 == 0 ? arg+1 : arg;
        }
    }
private:
    static double ewaldError(double width, double alpha, double target, int kmax) {
        const double temp = kmax*3.14159265358979323846/(width*alpha);
        return target-0.05*std::sqrt(width*alpha)*kmax*std::exp(-temp*temp);
    }
    const NonbondedForce& owner;
};
}
#endif
