/* STAND-IN for openmm/OpenMMException.h (compile check only). */
#ifndef MOCK_OPENMM_EXCEPTION_H_
#define MOCK_OPENMM_EXCEPTION_H_
#include <exception>
#include <string>
namespace OpenMM {
class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string& message) : message(message) {}
    ~OpenMMException() throw() {}
    const char* what() const throw() { return message.c_str(); }
private:
    std::string message;
};
}
#endif
