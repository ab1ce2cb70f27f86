/* STAND-IN for OpenMM's openmm/Vec3.h -- compile check of plugin/ only (OpenMM is not installed here). */
#ifndef MOCK_OPENMM_VEC3_H_
#define MOCK_OPENMM_VEC3_H_
namespace OpenMM {
class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    Vec3& operator+=(const Vec3& o) { data[0] += o[0]; data[1] += o[1]; data[2] += o[2]; return *this; }
private:
    double data[3];
};
}
#endif
