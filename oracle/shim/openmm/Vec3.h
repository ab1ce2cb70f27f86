// Minimal stand-in for OpenMM's openmm/Vec3.h (OpenMM is not vendored in the
// reference tree and not installed here).  Only what the three reference
// source files compiled by oracle/Makefile use: element access, + - * / and dot
// (call sites: ReferencePME.cpp:186-194,245; ReferenceSlicedLJCoulombIxn.cpp:229-234).
// Test infrastructure only -- never linked into the product library.
#ifndef SNB_SHIM_OPENMM_VEC3_H_
#define SNB_SHIM_OPENMM_VEC3_H_
#include <cassert>
#include <cmath>
namespace OpenMM {
class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    Vec3 operator+(const Vec3& r) const { return Vec3(data[0]+r[0], data[1]+r[1], data[2]+r[2]); }
    Vec3 operator-(const Vec3& r) const { return Vec3(data[0]-r[0], data[1]-r[1], data[2]-r[2]); }
    Vec3 operator-() const { return Vec3(-data[0], -data[1], -data[2]); }
    Vec3 operator*(double s) const { return Vec3(data[0]*s, data[1]*s, data[2]*s); }
    Vec3 operator/(double s) const { return Vec3(data[0]/s, data[1]/s, data[2]/s); }
    Vec3& operator+=(const Vec3& r) { data[0] += r[0]; data[1] += r[1]; data[2] += r[2]; return *this; }
    Vec3& operator-=(const Vec3& r) { data[0] -= r[0]; data[1] -= r[1]; data[2] -= r[2]; return *this; }
    Vec3& operator*=(double s) { data[0] *= s; data[1] *= s; data[2] *= s; return *this; }
    double dot(const Vec3& r) const { return data[0]*r[0] + data[1]*r[1] + data[2]*r[2]; }
private:
    double data[3];
};
static inline Vec3 operator*(double s, const Vec3& v) { return v*s; }
}
#endif
