/* STAND-IN for openmm/Force.h: what SlicedNonbondedForce inherits through NonbondedForce, with OpenMM's names and
 * signatures.  Enough to COMPILE the plugin against and, with src/mock_openmm.cpp, to LINK and RUN it (plugin/Makefile). */
#ifndef MOCK_OPENMM_FORCE_H_
#define MOCK_OPENMM_FORCE_H_
#include <string>
#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT
#endif
namespace OpenMM {
class ForceImpl;
class Context;
class ContextImpl;
class Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int group) { forceGroup = group; }
    const std::string& getName() const { return name; }
    void setName(const std::string& newName) { name = newName; }
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
protected:
    friend class Context;
    virtual ForceImpl* createImpl() const = 0;
    ForceImpl& getImplInContext(Context& context);
    const ForceImpl& getImplInContext(const Context& context) const;
    ContextImpl& getContextImpl(Context& context);
private:
    int forceGroup;
    std::string name;
};
}
#endif
