/* STAND-IN for openmm/internal/NonbondedForceImpl.h: the base the reference's SlicedNonbondedForceImpl derives from (it
 * overrides everything but the default parameters: NonbondedForceImpl::getDefaultParameters = the force's global parameters). */
#ifndef MOCK_OPENMM_NONBONDEDFORCEIMPL_H_
#define MOCK_OPENMM_NONBONDEDFORCEIMPL_H_
#include "ForceImpl.h"
#include "../NonbondedForce.h"
#include "../System.h"
#include "ContextImpl.h"
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
private:
    const NonbondedForce& owner;
};
}
#endif
