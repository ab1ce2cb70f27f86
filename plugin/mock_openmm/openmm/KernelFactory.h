/* STAND-IN for openmm/KernelFactory.h (compile check only). */
#ifndef MOCK_OPENMM_KERNELFACTORY_H_
#define MOCK_OPENMM_KERNELFACTORY_H_
#include "KernelImpl.h"
namespace OpenMM {
class ContextImpl;
class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
};
}
#endif
