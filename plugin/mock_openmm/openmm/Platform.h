/* STAND-IN for openmm/Platform.h: the registry a plugin's registerKernelFactories() works on and Platform::createKernel
 * (bodies in src/mock_openmm.cpp). */
#ifndef MOCK_OPENMM_PLATFORM_H_
#define MOCK_OPENMM_PLATFORM_H_
#include <map>
#include <string>
#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT
#endif
namespace OpenMM {
class Kernel;
class KernelFactory;
class KernelImpl;
class ContextImpl;
class Platform {
public:
    virtual ~Platform() {}
    static int getNumPlatforms();
    static Platform& getPlatform(int index);
    static Platform& getPlatformByName(const std::string& name);
    static void registerPlatform(Platform* platform);
    virtual const std::string& getName() const;
    const std::string& getPropertyDefaultValue(const std::string& property) const;
    void registerKernelFactory(const std::string& name, KernelFactory* factory);
    Kernel createKernel(const std::string& name, ContextImpl& context) const;
    KernelImpl* createKernelImpl(const std::string& name, ContextImpl& context) const;
private:
    std::map<std::string, KernelFactory*> factories;
};
}
#endif
