/* STAND-IN for openmm/internal/AssertionUtilities.h. */
#ifndef MOCK_OPENMM_ASSERTIONUTILITIES_H_
#define MOCK_OPENMM_ASSERTIONUTILITIES_H_
#include <sstream>
#include <string>
namespace OpenMM { void throwException(const char* file, int line, const std::string& details); }
#endif
