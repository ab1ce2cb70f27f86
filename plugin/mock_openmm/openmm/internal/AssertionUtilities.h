/* STAND-IN for openmm/internal/AssertionUtilities.h: OpenMM's assertion macros (names and meaning as in OpenMM 8.1). */
#ifndef MOCK_OPENMM_ASSERTIONUTILITIES_H_
#define MOCK_OPENMM_ASSERTIONUTILITIES_H_
#include <cmath>
#include <sstream>
#include <string>
namespace OpenMM { void throwException(const char* file, int line, const std::string& details); }
#define ASSERT(cond) {if (!(cond)) OpenMM::throwException(__FILE__, __LINE__, "");};
#define ASSERT_EQUAL(expected, found) {if (!((expected) == (found))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_TOL(expected, found, tol) {double _scale_ = std::abs(expected) > 1.0 ? std::abs(expected) : 1.0; if (!(std::abs((expected)-(found))/_scale_ <= (tol))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_VEC(expected, found, tol) {double _norm_ = std::sqrt((expected).dot(expected)); double _scale_ = _norm_ > 1.0 ? _norm_ : 1.0; if ((std::abs(((expected)[0])-((found)[0]))/_scale_ > (tol)) || (std::abs(((expected)[1])-((found)[1]))/_scale_ > (tol)) || (std::abs(((expected)[2])-((found)[2]))/_scale_ > (tol))) {std::stringstream details; details << " Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_USUALLY_EQUAL_TOL(expected, found, tol) ASSERT_EQUAL_TOL(expected, found, tol)
#define ASSERT_VALID_INDEX(index, vector) {if (index < 0 || index >= (int) vector.size()) OpenMM::throwException(__FILE__, __LINE__, "Index out of range");};
#endif
