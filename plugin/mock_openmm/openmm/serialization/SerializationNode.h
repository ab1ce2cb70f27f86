/* STAND-IN for openmm/serialization/SerializationNode.h (OpenMM 8.1's names and meaning): a named node with string
 * properties -- integers, booleans (stored as 0 / 1) and doubles (shortest decimal form that reads back to the same bits) are
 * strings underneath -- and child nodes.  Header-only. */
#ifndef MOCK_OPENMM_SERIALIZATIONNODE_H_
#define MOCK_OPENMM_SERIALIZATIONNODE_H_
#include "../OpenMMException.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class SerializationNode {
public:
    const std::string& getName() const { return name; }
    void setName(const std::string& n) { name = n; }
    const std::vector<SerializationNode>& getChildren() const { return children; }
    std::vector<SerializationNode>& getChildren() { return children; }
    const SerializationNode& getChildNode(const std::string& n) const {
        for (size_t i = 0; i < children.size(); i++)
            if (children[i].name == n)
                return children[i];
        throw OpenMMException("Unknown child '"+n+"' in node '"+name+"'");
    }
    SerializationNode& createChildNode(const std::string& n) {
        children.push_back(SerializationNode());
        children.back().setName(n);
        return children.back();
    }
    const std::map<std::string, std::string>& getProperties() const { return properties; }
    bool hasProperty(const std::string& n) const { return properties.find(n) != properties.end(); }
    const std::string& getStringProperty(const std::string& n) const {
        std::map<std::string, std::string>::const_iterator p = properties.find(n);
        if (p == properties.end())
            throw OpenMMException("Unknown property '"+n+"' in node '"+name+"'");
        return p->second;
    }
    const std::string& getStringProperty(const std::string& n, const std::string& defaultValue) const {
        std::map<std::string, std::string>::const_iterator p = properties.find(n);
        return p == properties.end() ? defaultValue : p->second;
    }
    SerializationNode& setStringProperty(const std::string& n, const std::string& value) { properties[n] = value; return *this; }
    int getIntProperty(const std::string& n) const { return atoi(getStringProperty(n).c_str()); }
    int getIntProperty(const std::string& n, int defaultValue) const { return hasProperty(n) ? getIntProperty(n) : defaultValue; }
    SerializationNode& setIntProperty(const std::string& n, int value) {
        char buffer[32];
        snprintf(buffer, sizeof(buffer), "%d", value);
        return setStringProperty(n, buffer);
    }
    bool getBoolProperty(const std::string& n) const { return getIntProperty(n) != 0; }
    bool getBoolProperty(const std::string& n, bool defaultValue) const { return hasProperty(n) ? getBoolProperty(n) : defaultValue; }
    SerializationNode& setBoolProperty(const std::string& n, bool value) { return setIntProperty(n, value ? 1 : 0); }
    double getDoubleProperty(const std::string& n) const { return strtod(getStringProperty(n).c_str(), NULL); }
    double getDoubleProperty(const std::string& n, double defaultValue) const { return hasProperty(n) ? getDoubleProperty(n) : defaultValue; }
    SerializationNode& setDoubleProperty(const std::string& n, double value) {
        /* the shortest decimal form that reads back to the same bits, written the way OpenMM's g_fmt writes it: plain
         * notation without a leading zero for moderate exponents, d[.ddd]e[-]xx beyond */
        char buffer[64];
        if (value != value)
            return setStringProperty(n, "nan");
        if (value == 0.0)
            return setStringProperty(n, std::signbit(value) ? "-0" : "0");
        if (std::isinf(value))
            return setStringProperty(n, value > 0 ? "inf" : "-inf");
        int digits = 1;
        for (; digits <= 17; digits++) {
            snprintf(buffer, sizeof(buffer), "%.*e", digits-1, value);
            if (strtod(buffer, NULL) == value)
                break;
        }
        const int exponent = atoi(strchr(buffer, 'e')+1);
        std::string text;
        if (exponent >= -4 && exponent < 16) {
            const int decimals = digits-1-exponent > 0 ? digits-1-exponent : 0;
            snprintf(buffer, sizeof(buffer), "%.*f", decimals, value);
            text = buffer;
            if (text.compare(0, 2, "0.") == 0) text.erase(0, 1);
            else if (text.compare(0, 3, "-0.") == 0) text.erase(1, 1);
        }
        else {
            /* mantissa digits as found, exponent with at least two digits (1e-07, 1.5e+300) */
            std::string mantissa(buffer, strchr(buffer, 'e')-buffer);
            snprintf(buffer, sizeof(buffer), "e%c%02d", exponent < 0 ? '-' : '+', exponent < 0 ? -exponent : exponent);
            text = mantissa+buffer;
        }
        return setStringProperty(n, text);
    }
private:
    std::string name;
    std::vector<SerializationNode> children;
    std::map<std::string, std::string> properties;
};
}
#endif
