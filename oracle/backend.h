/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the SlicedNonbondedForce path.
 * Nothing under oracle/ may be imported, linked or executed by the product
 * (openmm-lab_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker.
 *
 * backend.h: the arithmetic interface between the oracle harness
 * (oracle/harness.cpp, a restatement of the reference kernel glue) and one of
 * two interchangeable arithmetic back ends:
 *   - oracle/ref_backend.cpp : calls the reference's OWN sources, compiled
 *     unchanged from /root/reference (only built into oracle/_ref/);
 *   - oracle/port_backend.cpp: a from-scratch restatement of the same maths.
 */
#ifndef SNB_ORACLE_BACKEND_H_
#define SNB_ORACLE_BACKEND_H_

#include <set>
#include <utility>
#include <vector>

namespace snb_oracle {

struct PairSettings {
    int method;                 // SNB_NO_CUTOFF .. SNB_LJPME
    double cutoff;
    bool useSwitch;
    double switchingDistance;
    double rfDielectric;
    double box[3][3];           // rows a, b, c
    bool periodicExceptions;
    double alpha;               // Ewald / PME
    int grid[3];
    double dispersionAlpha;     // LJPME
    int dispersionGrid[3];
    int kmax[3];                // Ewald
};

typedef std::vector<std::pair<int, int> > PairList;

// ReferenceSlicedLJCoulombIxn::calculatePairIxn
// (platforms/reference/src/ReferenceSlicedLJCoulombIxn.cpp:512-536).
//   pos     [N][3]      params [N][3] = (sigma/2, 2*sqrt(eps), q)
//   lambdas [L][2] (Coul, vdW)          forces [N][3] added to
//   sliceEnergies [L][2] added to
void backendPairIxn(const PairSettings& s, int numAtoms, const double* pos, int numSubsets,
                    const int* subsets, const double* params, const double* lambdas,
                    const std::vector<std::set<int> >& exclusions, const PairList& neighbors,
                    double* forces, double* sliceEnergies, bool includeDirect, bool includeReciprocal);

// ReferenceSlicedLJCoulomb14::calculateBondIxn
// (platforms/reference/src/ReferenceSlicedLJCoulomb14.cpp:61-95).
//   params = (sigma, 4*eps, chargeProd); lambdas/sliceEnergies = the slice's [2]
void backend14Ixn(const PairSettings& s, bool periodic, int i, int j, const double* pos,
                  const double* params, double* forces, const double* lambdas, double* sliceEnergies);

const char* backendName();

}  // namespace snb_oracle

#endif
