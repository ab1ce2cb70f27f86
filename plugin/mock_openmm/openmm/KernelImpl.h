/* STAND-IN for openmm/KernelImpl.h (compile check only). */
#ifndef MOCK_OPENMM_KERNELIMPL_H_
#define MOCK_OPENMM_KERNELIMPL_H_
#include "Platform.h"
#include <string>
namespace OpenMM {
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
private:
    std::string name;
    const Platform* platform;
};
}
#endif
