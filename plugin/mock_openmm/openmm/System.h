/* STAND-IN for openmm/System.h (header-only). */
#ifndef MOCK_OPENMM_SYSTEM_H_
#define MOCK_OPENMM_SYSTEM_H_
#include "Force.h"
#include "Vec3.h"
#include <vector>
namespace OpenMM {
class System {
public:
    System() { box[0] = Vec3(2, 0, 0); box[1] = Vec3(0, 2, 0); box[2] = Vec3(0, 0, 2); }
    ~System() {
        for (size_t i = 0; i < forces.size(); i++)
            delete forces[i];
    }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size()-1; }
    int getNumParticles() const { return (int) masses.size(); }
    int addConstraint(int, int, double) { return numConstraints++; }
    int addForce(Force* force) { forces.push_back(force); return (int) forces.size()-1; }       /* the System owns its forces */
    int getNumForces() const { return (int) forces.size(); }
    Force& getForce(int index) { return *forces.at(index); }
    const Force& getForce(int index) const { return *forces.at(index); }
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
    bool usesPeriodicBoundaryConditions() const {
        for (size_t i = 0; i < forces.size(); i++)
            if (forces[i]->usesPeriodicBoundaryConditions())
                return true;
        return false;
    }
private:
    System(const System&);
    System& operator=(const System&);
    std::vector<double> masses;
    std::vector<Force*> forces;
    Vec3 box[3];
    int numConstraints = 0;
};
}
#endif
