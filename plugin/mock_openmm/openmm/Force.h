/* STAND-IN for openmm/Force.h (compile check only). */
#ifndef MOCK_OPENMM_FORCE_H_
#define MOCK_OPENMM_FORCE_H_
#include <string>
#define OPENMM_EXPORT
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
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
protected:
    virtual ForceImpl* createImpl() const = 0;
private:
    int forceGroup;
};
}
#endif
