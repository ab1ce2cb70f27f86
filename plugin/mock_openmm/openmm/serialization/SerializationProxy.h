/* STAND-IN for openmm/serialization/SerializationProxy.h: the registry of proxies by type and by type name.  Header-only. */
#ifndef MOCK_OPENMM_SERIALIZATIONPROXY_H_
#define MOCK_OPENMM_SERIALIZATIONPROXY_H_
#include "SerializationNode.h"
#include <map>
#include <string>
#include <typeinfo>
namespace OpenMM {
class SerializationProxy {
public:
    SerializationProxy(const std::string& typeName) : typeName(typeName) {}
    virtual ~SerializationProxy() {}
    const std::string& getTypeName() const { return typeName; }
    virtual void serialize(const void* object, SerializationNode& node) const = 0;
    virtual void* deserialize(const SerializationNode& node) const = 0;
    static void registerProxy(const std::type_info& type, const SerializationProxy* proxy) {
        byType()[type.name()] = proxy;
        byName()[proxy->getTypeName()] = proxy;
    }
    static const SerializationProxy& getProxy(const std::string& typeName) {
        std::map<std::string, const SerializationProxy*>::const_iterator p = byName().find(typeName);
        if (p == byName().end())
            throw OpenMMException("There is no serialization proxy registered for type "+typeName);
        return *p->second;
    }
    static const SerializationProxy& getProxy(const std::type_info& type) {
        std::map<std::string, const SerializationProxy*>::const_iterator p = byType().find(type.name());
        if (p == byType().end())
            throw OpenMMException(std::string("There is no serialization proxy registered for type ")+type.name());
        return *p->second;
    }
private:
    static std::map<std::string, const SerializationProxy*>& byType() { static std::map<std::string, const SerializationProxy*> m; return m; }
    static std::map<std::string, const SerializationProxy*>& byName() { static std::map<std::string, const SerializationProxy*> m; return m; }
    std::string typeName;
};
}
#endif
