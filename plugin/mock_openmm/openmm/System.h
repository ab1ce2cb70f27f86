/* STAND-IN for openmm/System.h (what the kernels read of it; header-only). */
#ifndef MOCK_OPENMM_SYSTEM_H_
#define MOCK_OPENMM_SYSTEM_H_
#include "Vec3.h"
#include <vector>
namespace OpenMM {
class System {
public:
    System() { box[0] = Vec3(2, 0, 0); box[1] = Vec3(0, 2, 0); box[2] = Vec3(0, 0, 2); }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size()-1; }
    int getNumParticles() const { return (int) masses.size(); }
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
private:
    std::vector<double> masses;
    Vec3 box[3];
};
}
#endif
