/* STAND-IN for openmm/Platform.h (compile check only). */
#ifndef MOCK_OPENMM_PLATFORM_H_
#define MOCK_OPENMM_PLATFORM_H_
#include <string>
namespace OpenMM {
class KernelFactory;
class Platform {
public:
    virtual ~Platform() {}
    static int getNumPlatforms();
    static Platform& getPlatform(int index);
    static Platform& getPlatformByName(const std::string& name);
    const std::string& getName() const;
    void registerKernelFactory(const std::string& name, KernelFactory* factory);
};
}
#endif
