/* STAND-IN for openmm/reference/ReferenceNeighborList.h: the list the reference's pair loops iterate, and OpenMM's
 * computeNeighborListVoxelHash restated with its SET semantics (every pair i < j that is not excluded and whose -- periodic:
 * minimum-image, reduced triclinic cell -- distance lies in [minDistance, maxDistance]) as a plain double loop: the
 * reference's tests are at most 1 200 particles. */
#ifndef MOCK_OPENMM_REFERENCENEIGHBORLIST_H_
#define MOCK_OPENMM_REFERENCENEIGHBORLIST_H_
#include "ReferenceForce.h"
#include <set>
#include <utility>
#include <vector>
namespace OpenMM {
typedef std::pair<int, int> AtomPair;
class NeighborList : public std::vector<AtomPair> {};
static inline void computeNeighborListVoxelHash(NeighborList& neighborList, int nAtoms, const std::vector<Vec3>& atomLocations,
        const std::vector<std::set<int> >& exclusions, const Vec3* periodicBoxVectors, bool usePeriodic, double maxDistance,
        double minDistance = 0.0, bool reportSymmetricPairs = false) {
    neighborList.clear();
    const double max2 = maxDistance*maxDistance, min2 = minDistance*minDistance;
    for (int i = 0; i < nAtoms; i++)
        for (int j = i+1; j < nAtoms; j++) {
            if (exclusions[i].find(j) != exclusions[i].end())
                continue;
            double delta[ReferenceForce::LastDeltaRIndex];
            if (usePeriodic)
                ReferenceForce::getDeltaRPeriodic(atomLocations[i], atomLocations[j], periodicBoxVectors, delta);
            else
                ReferenceForce::getDeltaR(atomLocations[i], atomLocations[j], delta);
            const double r2 = delta[ReferenceForce::R2Index];
            if (r2 > max2 || r2 < min2)
                continue;
            neighborList.push_back(AtomPair(i, j));
            if (reportSymmetricPairs)
                neighborList.push_back(AtomPair(j, i));
        }
}
}
#endif
