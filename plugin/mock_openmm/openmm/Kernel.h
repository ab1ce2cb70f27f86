/* STAND-IN for openmm/Kernel.h: the reference-counted handle to a KernelImpl. */
#ifndef MOCK_OPENMM_KERNEL_H_
#define MOCK_OPENMM_KERNEL_H_
#include "KernelImpl.h"
#include <memory>
namespace OpenMM {
class Kernel {
public:
    Kernel() {}
    Kernel(KernelImpl* impl) : impl(impl) {}
    std::string getName() const { return impl->getName(); }
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
    template <class T> const T& getAs() const { return dynamic_cast<const T&>(*impl); }
private:
    std::shared_ptr<KernelImpl> impl;
};
}
#endif
