/* STAND-IN for openmm/internal/ForceImpl.h. */
#ifndef MOCK_OPENMM_FORCEIMPL_H_
#define MOCK_OPENMM_FORCEIMPL_H_
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class Force;
class ContextImpl;
class ForceImpl {
public:
    virtual ~ForceImpl() {}
    virtual void initialize(ContextImpl& context) = 0;
    virtual const Force& getOwner() const = 0;
    virtual void updateContextState(ContextImpl&, bool&) {}
    virtual double calcForcesAndEnergy(ContextImpl& context, bool includeForces, bool includeEnergy, int groups) = 0;
    virtual std::map<std::string, double> getDefaultParameters() { return std::map<std::string, double>(); }
    virtual std::vector<std::string> getKernelNames() = 0;
};
}
#endif
