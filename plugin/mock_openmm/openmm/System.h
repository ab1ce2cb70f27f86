/* STAND-IN for openmm/System.h (compile check only). */
#ifndef MOCK_OPENMM_SYSTEM_H_
#define MOCK_OPENMM_SYSTEM_H_
#include "Vec3.h"
namespace OpenMM {
class System {
public:
    int getNumParticles() const;
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const;
};
}
#endif
