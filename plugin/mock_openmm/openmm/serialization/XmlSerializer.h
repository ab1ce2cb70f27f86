/* STAND-IN for openmm/serialization/XmlSerializer.h: writes / reads a SerializationNode tree the way OpenMM 8.1 does
 * (root element named by the caller, carrying `type` and `openmmVersion`; properties as attributes in std::map order;
 * children indented with tabs) -- the same layout openmm-lab_b200/serialization.py writes and reads.  Header-only. */
#ifndef MOCK_OPENMM_XMLSERIALIZER_H_
#define MOCK_OPENMM_XMLSERIALIZER_H_
#include "SerializationNode.h"
#include "SerializationProxy.h"
#include <cctype>
#include <istream>
#include <iterator>
#include <ostream>
#include <string>
#include <typeinfo>
namespace OpenMM {
class XmlSerializer {
public:
    template <class T> static void serialize(const T* object, const std::string& rootName, std::ostream& stream) {
        const SerializationProxy& proxy = SerializationProxy::getProxy(typeid(*object));
        SerializationNode node;
        node.setName(rootName);
        proxy.serialize(object, node);
        if (node.hasProperty("type"))
            throw OpenMMException(proxy.getTypeName()+" created node with reserved property 'type'");
        node.setStringProperty("type", proxy.getTypeName());
        node.setStringProperty("openmmVersion", "8.1");
        stream << "<?xml version=\"1.0\" ?>\n";
        write(node, stream, 0);
    }
    template <class T> static T* deserialize(std::istream& stream) {
        const std::string text((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
        size_t pos = 0;
        SerializationNode root = parseElement(text, pos);
        const std::string type = root.getStringProperty("type");
        return reinterpret_cast<T*>(SerializationProxy::getProxy(type).deserialize(root));
    }
private:
    static std::string escape(const std::string& s) {
        std::string out;
        for (size_t i = 0; i < s.size(); i++)
            switch (s[i]) {
                case '&': out += "&amp;"; break;
                case '<': out += "&lt;"; break;
                case '>': out += "&gt;"; break;
                case '"': out += "&quot;"; break;
                default: out += s[i];
            }
        return out;
    }
    static std::string unescape(const std::string& s) {
        std::string out;
        for (size_t i = 0; i < s.size(); i++) {
            if (s.compare(i, 5, "&amp;") == 0) { out += '&'; i += 4; }
            else if (s.compare(i, 4, "&lt;") == 0) { out += '<'; i += 3; }
            else if (s.compare(i, 4, "&gt;") == 0) { out += '>'; i += 3; }
            else if (s.compare(i, 6, "&quot;") == 0) { out += '"'; i += 5; }
            else if (s.compare(i, 6, "&apos;") == 0) { out += '\''; i += 5; }
            else out += s[i];
        }
        return out;
    }
    static void write(const SerializationNode& node, std::ostream& stream, int depth) {
        stream << std::string(depth, '\t') << '<' << node.getName();
        for (std::map<std::string, std::string>::const_iterator p = node.getProperties().begin(); p != node.getProperties().end(); ++p)
            stream << ' ' << p->first << "=\"" << escape(p->second) << '"';
        if (node.getChildren().empty()) {
            stream << "/>\n";
            return;
        }
        stream << ">\n";
        for (size_t i = 0; i < node.getChildren().size(); i++)
            write(node.getChildren()[i], stream, depth+1);
        stream << std::string(depth, '\t') << "</" << node.getName() << ">\n";
    }
    static void skipSpace(const std::string& t, size_t& pos) {
        while (pos < t.size() && isspace((unsigned char) t[pos]))
            pos++;
    }
    static void skipProlog(const std::string& t, size_t& pos) {
        for (;;) {
            skipSpace(t, pos);
            if (t.compare(pos, 2, "<?") == 0) pos = t.find("?>", pos)+2;
            else if (t.compare(pos, 4, "<!--") == 0) pos = t.find("-->", pos)+3;
            else return;
        }
    }
    static SerializationNode parseElement(const std::string& t, size_t& pos) {
        skipProlog(t, pos);
        if (pos >= t.size() || t[pos] != '<')
            throw OpenMMException("Error parsing XML: expected an element");
        pos++;
        size_t start = pos;
        while (pos < t.size() && !isspace((unsigned char) t[pos]) && t[pos] != '>' && t[pos] != '/')
            pos++;
        SerializationNode node;
        node.setName(t.substr(start, pos-start));
        for (;;) {
            skipSpace(t, pos);
            if (pos >= t.size())
                throw OpenMMException("Error parsing XML: unterminated element");
            if (t[pos] == '/') {            /* <name ... /> */
                pos += 2;
                return node;
            }
            if (t[pos] == '>') {
                pos++;
                break;
            }
            start = pos;
            while (pos < t.size() && t[pos] != '=' && !isspace((unsigned char) t[pos]))
                pos++;
            const std::string key = t.substr(start, pos-start);
            skipSpace(t, pos);
            if (t[pos] != '=')
                throw OpenMMException("Error parsing XML: attribute without a value");
            pos++;
            skipSpace(t, pos);
            const char quote = t[pos++];
            start = pos;
            pos = t.find(quote, pos);
            if (pos == std::string::npos)
                throw OpenMMException("Error parsing XML: unterminated attribute");
            node.setStringProperty(key, unescape(t.substr(start, pos-start)));
            pos++;
        }
        for (;;) {                          /* children until </name> */
            skipProlog(t, pos);
            if (t.compare(pos, 2, "</") == 0) {
                pos = t.find('>', pos)+1;
                return node;
            }
            if (pos >= t.size())
                throw OpenMMException("Error parsing XML: missing end tag");
            if (t[pos] != '<') {            /* text content is not part of the format */
                pos = t.find('<', pos);
                continue;
            }
            node.getChildren().push_back(parseElement(t, pos));
        }
    }
};
}
#endif
