// Stand-in for openmm/OpenMMException.h (used at ReferenceSlicedLJCoulombIxn.cpp:190,192,254).
#ifndef SNB_SHIM_OPENMM_EXCEPTION_H_
#define SNB_SHIM_OPENMM_EXCEPTION_H_
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
