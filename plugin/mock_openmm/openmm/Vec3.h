/* STAND-IN for OpenMM's openmm/Vec3.h (OpenMM is not installed here): the operators the kernels and the reference's tests use. */
#ifndef MOCK_OPENMM_VEC3_H_
#define MOCK_OPENMM_VEC3_H_
#include <ostream>
namespace OpenMM {
class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    bool operator==(const Vec3& o) const { return data[0] == o[0] && data[1] == o[1] && data[2] == o[2]; }
    bool operator!=(const Vec3& o) const { return !(*this == o); }
    Vec3 operator+() const { return *this; }
    Vec3 operator-() const { return Vec3(-data[0], -data[1], -data[2]); }
    Vec3 operator+(const Vec3& o) const { return Vec3(data[0]+o[0], data[1]+o[1], data[2]+o[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(data[0]-o[0], data[1]-o[1], data[2]-o[2]); }
    Vec3& operator+=(const Vec3& o) { data[0] += o[0]; data[1] += o[1]; data[2] += o[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { data[0] -= o[0]; data[1] -= o[1]; data[2] -= o[2]; return *this; }
    Vec3 operator*(double s) const { return Vec3(data[0]*s, data[1]*s, data[2]*s); }
    Vec3& operator*=(double s) { data[0] *= s; data[1] *= s; data[2] *= s; return *this; }
    Vec3 operator/(double s) const { return Vec3(data[0]/s, data[1]/s, data[2]/s); }
    Vec3& operator/=(double s) { data[0] /= s; data[1] /= s; data[2] /= s; return *this; }
    double dot(const Vec3& o) const { return data[0]*o[0]+data[1]*o[1]+data[2]*o[2]; }
    Vec3 cross(const Vec3& o) const { return Vec3(data[1]*o[2]-data[2]*o[1], data[2]*o[0]-data[0]*o[2], data[0]*o[1]-data[1]*o[0]); }
private:
    double data[3];
};
static inline Vec3 operator*(double s, const Vec3& v) { return v*s; }
template <class CHAR, class TRAITS>
std::basic_ostream<CHAR, TRAITS>& operator<<(std::basic_ostream<CHAR, TRAITS>& o, const Vec3& v) { o << '[' << v[0] << ", " << v[1] << ", " << v[2] << ']'; return o; }
}
#endif
