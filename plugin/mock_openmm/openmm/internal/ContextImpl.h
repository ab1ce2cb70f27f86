/* STAND-IN for openmm/internal/ContextImpl.h (compile check only). */
#ifndef MOCK_OPENMM_CONTEXTIMPL_H_
#define MOCK_OPENMM_CONTEXTIMPL_H_
#include "../System.h"
#include <string>
namespace OpenMM {
class ContextImpl {
public:
    void* getPlatformData();
    double getParameter(std::string name);
    const System& getSystem() const;
};
}
#endif
