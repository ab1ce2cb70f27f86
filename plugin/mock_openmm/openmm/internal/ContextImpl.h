/* STAND-IN for openmm/internal/ContextImpl.h: the three calls the kernels make on it (header-only). */
#ifndef MOCK_OPENMM_CONTEXTIMPL_H_
#define MOCK_OPENMM_CONTEXTIMPL_H_
#include "../OpenMMException.h"
#include "../System.h"
#include <map>
#include <string>
namespace OpenMM {
class ContextImpl {
public:
    ContextImpl(const System& system, void* platformData) : system(system), platformData(platformData) {}
    void* getPlatformData() { return platformData; }
    double getParameter(std::string name) {
        if (parameters.find(name) == parameters.end())
            throw OpenMMException("Called getParameter() with invalid parameter name");
        return parameters[name];
    }
    void setParameter(std::string name, double value) { parameters[name] = value; }
    const System& getSystem() const { return system; }
private:
    const System& system;
    void* platformData;
    std::map<std::string, double> parameters;
};
}
#endif
