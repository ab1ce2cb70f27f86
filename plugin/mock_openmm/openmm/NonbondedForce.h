/* STAND-IN for openmm/NonbondedForce.h: the parameter container SlicedNonbondedForce inherits (OpenMM 8.1's names and
 * signatures; header-only, so that the reference's own openmmapi/src/SlicedNonbondedForce.cpp links against it). */
#ifndef MOCK_OPENMM_NONBONDEDFORCE_H_
#define MOCK_OPENMM_NONBONDEDFORCE_H_
#include "Force.h"
#include "OpenMMException.h"
#include <cmath>
#include <set>
#include <string>
#include <utility>
#include <vector>
namespace OpenMM {
class NonbondedForce : public Force {
public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2, Ewald = 3, PME = 4, LJPME = 5 };
    NonbondedForce() : method(NoCutoff), cutoff(1.0), useSwitch(false), switchingDistance(-1.0), rfDielectric(78.3), ewaldErrorTol(5e-4),
        alpha(0), dalpha(0), nx(0), ny(0), nz(0), dnx(0), dny(0), dnz(0), useDispersionCorrection(true), exceptionsUsePeriodic(false),
        recipForceGroup(-1), includeDirectSpace(true) {}
    NonbondedMethod getNonbondedMethod() const { return method; }
    void setNonbondedMethod(NonbondedMethod m) { method = m; }
    double getCutoffDistance() const { return cutoff; }
    void setCutoffDistance(double d) { cutoff = d; }
    bool getUseSwitchingFunction() const { return useSwitch; }
    void setUseSwitchingFunction(bool use) { useSwitch = use; }
    double getSwitchingDistance() const { return switchingDistance; }
    void setSwitchingDistance(double d) { switchingDistance = d; }
    double getReactionFieldDielectric() const { return rfDielectric; }
    void setReactionFieldDielectric(double d) { rfDielectric = d; }
    double getEwaldErrorTolerance() const { return ewaldErrorTol; }
    void setEwaldErrorTolerance(double tol) { ewaldErrorTol = tol; }
    void getPMEParameters(double& a, int& x, int& y, int& z) const { a = alpha; x = nx; y = ny; z = nz; }
    void setPMEParameters(double a, int x, int y, int z) { alpha = a; nx = x; ny = y; nz = z; }
    void getLJPMEParameters(double& a, int& x, int& y, int& z) const { a = dalpha; x = dnx; y = dny; z = dnz; }
    void setLJPMEParameters(double a, int x, int y, int z) { dalpha = a; dnx = x; dny = y; dnz = z; }
    bool getUseDispersionCorrection() const { return useDispersionCorrection; }
    void setUseDispersionCorrection(bool use) { useDispersionCorrection = use; }
    bool getExceptionsUsePeriodicBoundaryConditions() const { return exceptionsUsePeriodic; }
    void setExceptionsUsePeriodicBoundaryConditions(bool periodic) { exceptionsUsePeriodic = periodic; }
    int getReciprocalSpaceForceGroup() const { return recipForceGroup; }
    void setReciprocalSpaceForceGroup(int group) { recipForceGroup = group; }
    bool getIncludeDirectSpace() const { return includeDirectSpace; }
    void setIncludeDirectSpace(bool include) { includeDirectSpace = include; }
    bool usesPeriodicBoundaryConditions() const { return method == CutoffPeriodic || method == Ewald || method == PME || method == LJPME; }

    int getNumParticles() const { return (int) particles.size(); }
    int addParticle(double charge, double sigma, double epsilon) { particles.push_back(Particle(charge, sigma, epsilon)); return (int) particles.size()-1; }
    void getParticleParameters(int index, double& charge, double& sigma, double& epsilon) const {
        charge = particles.at(index).q; sigma = particles.at(index).sigma; epsilon = particles.at(index).epsilon;
    }
    void setParticleParameters(int index, double charge, double sigma, double epsilon) { particles.at(index) = Particle(charge, sigma, epsilon); }
    int getNumExceptions() const { return (int) exceptions.size(); }
    int addException(int particle1, int particle2, double chargeProd, double sigma, double epsilon, bool replace = false) {
        for (size_t i = 0; i < exceptions.size(); i++)
            if ((exceptions[i].p1 == particle1 && exceptions[i].p2 == particle2) || (exceptions[i].p1 == particle2 && exceptions[i].p2 == particle1)) {
                if (!replace)
                    throw OpenMMException("NonbondedForce: There is already an exception for these particles");
                exceptions[i] = Exception(particle1, particle2, chargeProd, sigma, epsilon);
                return (int) i;
            }
        exceptions.push_back(Exception(particle1, particle2, chargeProd, sigma, epsilon));
        return (int) exceptions.size()-1;
    }
    void getExceptionParameters(int index, int& particle1, int& particle2, double& chargeProd, double& sigma, double& epsilon) const {
        const Exception& e = exceptions.at(index);
        particle1 = e.p1; particle2 = e.p2; chargeProd = e.q; sigma = e.sigma; epsilon = e.epsilon;
    }
    void setExceptionParameters(int index, int particle1, int particle2, double chargeProd, double sigma, double epsilon) {
        exceptions.at(index) = Exception(particle1, particle2, chargeProd, sigma, epsilon);
    }
    int getNumGlobalParameters() const { return (int) globals.size(); }
    int addGlobalParameter(const std::string& name, double defaultValue) { globals.push_back(Global(name, defaultValue)); return (int) globals.size()-1; }
    const std::string& getGlobalParameterName(int index) const { return globals.at(index).name; }
    double getGlobalParameterDefaultValue(int index) const { return globals.at(index).value; }
    int getNumParticleParameterOffsets() const { return (int) particleOffsets.size(); }
    int addParticleParameterOffset(const std::string& parameter, int particleIndex, double chargeScale, double sigmaScale, double epsilonScale) {
        particleOffsets.push_back(Offset(parameter, particleIndex, chargeScale, sigmaScale, epsilonScale)); return (int) particleOffsets.size()-1;
    }
    void getParticleParameterOffset(int index, std::string& parameter, int& particleIndex, double& chargeScale, double& sigmaScale, double& epsilonScale) const {
        const Offset& o = particleOffsets.at(index);
        parameter = o.parameter; particleIndex = o.index; chargeScale = o.q; sigmaScale = o.sigma; epsilonScale = o.epsilon;
    }
    int getNumExceptionParameterOffsets() const { return (int) exceptionOffsets.size(); }
    int addExceptionParameterOffset(const std::string& parameter, int exceptionIndex, double chargeProdScale, double sigmaScale, double epsilonScale) {
        exceptionOffsets.push_back(Offset(parameter, exceptionIndex, chargeProdScale, sigmaScale, epsilonScale)); return (int) exceptionOffsets.size()-1;
    }
    void getExceptionParameterOffset(int index, std::string& parameter, int& exceptionIndex, double& chargeProdScale, double& sigmaScale, double& epsilonScale) const {
        const Offset& o = exceptionOffsets.at(index);
        parameter = o.parameter; exceptionIndex = o.index; chargeProdScale = o.q; sigmaScale = o.sigma; epsilonScale = o.epsilon;
    }
    /* OpenMM's NonbondedForce::createExceptionsFromBonds: 1-2 and 1-3 pairs excluded, 1-4 pairs scaled */
    void createExceptionsFromBonds(const std::vector<std::pair<int, int> >& bonds, double coulomb14Scale, double lj14Scale) {
        const int n = getNumParticles();
        std::vector<std::set<int> > bonded12(n);
        for (size_t i = 0; i < bonds.size(); i++) {
            bonded12[bonds[i].first].insert(bonds[i].second);
            bonded12[bonds[i].second].insert(bonds[i].first);
        }
        for (int i = 0; i < n; i++) {
            std::set<int> within2, within3;
            for (std::set<int>::const_iterator j = bonded12[i].begin(); j != bonded12[i].end(); ++j) {
                within2.insert(*j);
                within2.insert(bonded12[*j].begin(), bonded12[*j].end());
            }
            within2.erase(i);
            within3 = within2;
            for (std::set<int>::const_iterator j = within2.begin(); j != within2.end(); ++j)
                within3.insert(bonded12[*j].begin(), bonded12[*j].end());
            within3.erase(i);
            for (std::set<int>::const_iterator j = within3.begin(); j != within3.end(); ++j) {
                if (*j >= i)
                    continue;
                if (within2.find(*j) != within2.end())
                    addException(*j, i, 0.0, 1.0, 0.0);
                else {
                    const Particle& a = particles[*j];
                    const Particle& b = particles[i];
                    addException(*j, i, coulomb14Scale*a.q*b.q, 0.5*(a.sigma+b.sigma), lj14Scale*std::sqrt(a.epsilon*b.epsilon));
                }
            }
        }
    }
    /* (plain NonbondedForce in a stand-in Context: evaluated as a one-subset SlicedNonbondedForce; src/mock_openmm.cpp) */
    void updateParametersInContext(Context& context);
protected:
    ForceImpl* createImpl() const;
private:
    struct Particle { double q, sigma, epsilon; Particle(double q, double s, double e) : q(q), sigma(s), epsilon(e) {} };
    struct Exception { int p1, p2; double q, sigma, epsilon; Exception(int a, int b, double q, double s, double e) : p1(a), p2(b), q(q), sigma(s), epsilon(e) {} };
    struct Global { std::string name; double value; Global(const std::string& n, double v) : name(n), value(v) {} };
    struct Offset { std::string parameter; int index; double q, sigma, epsilon;
                    Offset(const std::string& p, int i, double q, double s, double e) : parameter(p), index(i), q(q), sigma(s), epsilon(e) {} };
    NonbondedMethod method;
    double cutoff;
    bool useSwitch;
    double switchingDistance, rfDielectric, ewaldErrorTol, alpha, dalpha;
    int nx, ny, nz, dnx, dny, dnz;
    bool useDispersionCorrection, exceptionsUsePeriodic;
    int recipForceGroup;
    bool includeDirectSpace;
    std::vector<Particle> particles;
    std::vector<Exception> exceptions;
    std::vector<Global> globals;
    std::vector<Offset> particleOffsets, exceptionOffsets;
};
}
#endif
