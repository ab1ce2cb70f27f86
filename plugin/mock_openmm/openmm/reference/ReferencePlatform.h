/* STAND-IN for openmm/reference/ReferencePlatform.h: the PlatformData fields the kernels use
 * (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:35-58). */
#ifndef MOCK_OPENMM_REFERENCEPLATFORM_H_
#define MOCK_OPENMM_REFERENCEPLATFORM_H_
#include "../Platform.h"
#include "../Vec3.h"
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class ReferencePlatform : public Platform {
public:
    class PlatformData {
    public:
        std::vector<Vec3>* positions;
        std::vector<Vec3>* velocities;
        std::vector<Vec3>* forces;
        Vec3* periodicBoxVectors;
        std::map<std::string, double>* energyParameterDerivatives;
    };
    const std::string& getName() const {
        static const std::string name = "Reference";
        return name;
    }
};
}
#endif
