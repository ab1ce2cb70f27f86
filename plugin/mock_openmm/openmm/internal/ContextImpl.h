/* STAND-IN for openmm/internal/ContextImpl.h: what kernels and ForceImpls call on it (header-only). */
#ifndef MOCK_OPENMM_CONTEXTIMPL_H_
#define MOCK_OPENMM_CONTEXTIMPL_H_
#include "../OpenMMException.h"
#include "../Platform.h"
#include "../System.h"
#include <map>
#include <string>
namespace OpenMM {
class ContextImpl {
public:
    ContextImpl(const System& system, void* platformData, Platform* platform = 0) : system(system), platformData(platformData), platform(platform) {}
    void* getPlatformData() { return platformData; }
    Platform& getPlatform() {
        if (platform == 0)
            throw OpenMMException("mock OpenMM: this ContextImpl has no Platform");
        return *platform;
    }
    double getParameter(std::string name) {
        if (parameters.find(name) == parameters.end())
            throw OpenMMException("Called getParameter() with invalid parameter name");
        return parameters[name];
    }
    void setParameter(std::string name, double value) { parameters[name] = value; }
    const std::map<std::string, double>& getParameters() const { return parameters; }
    const System& getSystem() const { return system; }
    void systemChanged() {}
    /* (what the reference's ExtendedCustomCVForce kernel, in the same source file as the nonbonded one, calls) */
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) { a = box[0]; b = box[1]; c = box[2]; }
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
    double getTime() const { return time; }
    void setTime(double t) { time = t; }
private:
    const System& system;
    void* platformData;
    Platform* platform;
    std::map<std::string, double> parameters;
    Vec3 box[3];
    double time = 0;
};
}
#endif
