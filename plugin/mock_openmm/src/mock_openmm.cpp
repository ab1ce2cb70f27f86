/* Bodies of the stand-in OpenMM classes that are not header-only: the platform registry and the exception helper.
 * Test infrastructure of plugin/Makefile's run-check (OpenMM is not installed in this image). */
#include "openmm/Force.h"
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"
#include "openmm/internal/AssertionUtilities.h"

#include <vector>

namespace OpenMM {

static std::vector<Platform*>& platforms() {
    static std::vector<Platform*> list;
    return list;
}

int Platform::getNumPlatforms() { return (int) platforms().size(); }
Platform& Platform::getPlatform(int index) { return *platforms().at(index); }
Platform& Platform::getPlatformByName(const std::string& name) {
    for (size_t i = 0; i < platforms().size(); i++)
        if (platforms()[i]->getName() == name)
            return *platforms()[i];
    throw OpenMMException("There is no registered Platform called \""+name+"\"");
}
void Platform::registerPlatform(Platform* platform) { platforms().push_back(platform); }
const std::string& Platform::getName() const {
    static const std::string name = "Platform";
    return name;
}
void Platform::registerKernelFactory(const std::string& name, KernelFactory* factory) {
    factories[name] = factory;      /* a later registration under the same kernel name replaces the earlier one */
}
KernelImpl* Platform::createKernelImpl(const std::string& name, ContextImpl& context) const {
    std::map<std::string, KernelFactory*>::const_iterator f = factories.find(name);
    if (f == factories.end())
        throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
    return f->second->createKernelImpl(name, *this, context);
}

void throwException(const char* file, int line, const std::string& details) {
    std::stringstream message;
    message << "Assertion failure at " << file << ":" << line << ".  " << details;
    throw OpenMMException(message.str());
}

ForceImpl& Force::getImplInContext(Context&) { throw OpenMMException("mock OpenMM: no Context"); }
const ForceImpl& Force::getImplInContext(const Context&) const { throw OpenMMException("mock OpenMM: no Context"); }
ContextImpl& Force::getContextImpl(Context&) { throw OpenMMException("mock OpenMM: no Context"); }

}
