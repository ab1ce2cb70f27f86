// Stand-in for openmm/reference/ReferenceForce.h: displacement helpers used at
// ReferenceSlicedLJCoulombIxn.cpp:359-362,444-450,556-570 and ReferenceSlicedLJCoulomb14.cpp:64-75.
// Semantics restated from OpenMM's public behaviour: delta = J - I; the periodic
// variant removes whole box vectors c, then b, then a (reduced triclinic cell).
#ifndef SNB_SHIM_REFERENCEFORCE_H_
#define SNB_SHIM_REFERENCEFORCE_H_
#include "openmm/Vec3.h"
#include <cmath>
namespace OpenMM {
class ReferenceForce {
public:
    static const int XIndex = 0;
    static const int YIndex = 1;
    static const int ZIndex = 2;
    static const int R2Index = 3;
    static const int RIndex = 4;
    static const int LastDeltaRIndex = 5;
    static void getDeltaR(const Vec3& atomI, const Vec3& atomJ, double* deltaR) {
        Vec3 d = atomJ - atomI;
        store(d, deltaR);
    }
    static void getDeltaRPeriodic(const Vec3& atomI, const Vec3& atomJ, const Vec3* box, double* deltaR) {
        Vec3 d = atomJ - atomI;
        d -= box[2]*std::floor(d[2]/box[2][2] + 0.5);
        d -= box[1]*std::floor(d[1]/box[1][1] + 0.5);
        d -= box[0]*std::floor(d[0]/box[0][0] + 0.5);
        store(d, deltaR);
    }
private:
    static void store(const Vec3& d, double* deltaR) {
        deltaR[XIndex] = d[0];
        deltaR[YIndex] = d[1];
        deltaR[ZIndex] = d[2];
        deltaR[R2Index] = d.dot(d);
        deltaR[RIndex] = std::sqrt(deltaR[R2Index]);
    }
};
}
#endif
