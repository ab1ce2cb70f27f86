/* STAND-IN for OpenMM 8.1's platforms/cuda headers (CudaContext.h, CudaArray.h, CudaForceInfo.h, CudaPlatform.h): the members
 * the binding of plugin/src/B200CudaOpenMMLabKernels.cpp uses, with OpenMM's names and signatures.  Compile check only. */
#ifndef MOCK_OPENMM_CUDACONTEXT_H_
#define MOCK_OPENMM_CUDACONTEXT_H_
#include "../Platform.h"
#include "../Vec3.h"
#include <cuda.h>
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class CudaArray {
public:
    CUdeviceptr& getDevicePointer();
};
class CudaForceInfo {
public:
    virtual ~CudaForceInfo() {}
    virtual bool areParticlesIdentical(int particle1, int particle2) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int index, std::vector<int>& particles) {}
    virtual bool areGroupsIdentical(int group1, int group2) { return true; }
};
class CudaContext {
public:
    void setAsCurrent();
    int getDeviceIndex() const;
    int getContextIndex() const;
    CUstream getCurrentStream();
    int getNumAtoms() const;
    int getPaddedNumAtoms() const;
    bool getUseDoublePrecision() const;
    bool getUseMixedPrecision() const;
    CudaArray& getPosq();
    CudaArray& getPosqCorrection();
    CudaArray& getAtomIndexArray();
    CudaArray& getForce();
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const;
    void addForce(CudaForceInfo* force);
    std::map<std::string, double>& getEnergyParamDerivWorkspace();
};
class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        std::vector<CudaContext*> contexts;
    };
    static const std::string& CudaDeterministicForces();
    const std::string& getPropertyValue(const class Context& context, const std::string& property) const;
};
}
#endif
