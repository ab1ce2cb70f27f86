/* STAND-IN for openmm/Context.h: a Context that owns one ForceImpl per Force of the System and sums their
 * calcForcesAndEnergy over the requested force groups -- what OpenMM's ContextImpl::calcForcesAndEnergy does -- on the
 * Reference platform's PlatformData.  Bodies in src/mock_openmm.cpp. */
#ifndef MOCK_OPENMM_CONTEXT_H_
#define MOCK_OPENMM_CONTEXT_H_
#include "Integrator.h"
#include "Platform.h"
#include "State.h"
#include "System.h"
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class ContextImpl;
class ForceImpl;
class Context {
public:
    Context(const System& system, Integrator& integrator, Platform& platform);
    Context(const System& system, Integrator& integrator, Platform& platform, const std::map<std::string, std::string>& properties);
    ~Context();
    const System& getSystem() const;
    Platform& getPlatform();
    void setPositions(const std::vector<Vec3>& positions);
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c);
    State getState(int types, bool enforcePeriodicBox = false, int groups = 0xFFFFFFFF) const;
    void setParameter(const std::string& name, double value);
    double getParameter(const std::string& name) const;
    const std::map<std::string, double>& getParameters() const;
    void reinitialize(bool preserveState = false);
private:
    friend class Force;
    struct Data;
    Data* data;
    void build();
    void destroy();
    Context(const Context&);
    Context& operator=(const Context&);
};
}
#endif
