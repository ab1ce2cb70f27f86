/*
 * TEST INFRASTRUCTURE ONLY.  Arithmetic back end that calls the reference's OWN
 * Reference-platform sources, compiled unchanged from /root/reference by
 * oracle/Makefile into oracle/_ref/ (never copied into this repository):
 *   platforms/reference/src/ReferenceSlicedLJCoulombIxn.cpp
 *   platforms/reference/src/ReferenceSlicedLJCoulomb14.cpp
 *   platforms/reference/src/ReferencePME.cpp
 * This file only converts flat arrays to the container types those classes take,
 * following the call sequence of ReferenceCalcSlicedNonbondedForceKernel::execute
 * (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:197-237).
 */
#include "backend.h"
#include "slicednb.h"
#include "internal/ReferenceSlicedLJCoulombIxn.h"
#include "internal/ReferenceSlicedLJCoulomb14.h"

namespace snb_oracle {

void backendPairIxn(const PairSettings& s, int numAtoms, const double* pos, int numSubsets,
                    const int* subsets, const double* params, const double* lambdas,
                    const std::vector<std::set<int> >& exclusions, const PairList& neighbors,
                    double* forces, double* sliceEnergies, bool includeDirect, bool includeReciprocal) {
    int numSlices = numSubsets*(numSubsets+1)/2;
    std::vector<OpenMM::Vec3> posData(numAtoms), forceData(numAtoms);
    std::vector<int> subsetData(subsets, subsets+numAtoms);
    std::vector<std::vector<double> > paramData(numAtoms, std::vector<double>(4, 0.0));
    for (int i = 0; i < numAtoms; i++) {
        posData[i] = OpenMM::Vec3(pos[3*i], pos[3*i+1], pos[3*i+2]);
        forceData[i] = OpenMM::Vec3(forces[3*i], forces[3*i+1], forces[3*i+2]);
        for (int k = 0; k < 3; k++)
            paramData[i][k] = params[3*i+k];
    }
    std::vector<std::vector<double> > lambdaData(numSlices, std::vector<double>(2)), energyData(numSlices, std::vector<double>(2));
    for (int i = 0; i < numSlices; i++)
        for (int t = 0; t < 2; t++) {
            lambdaData[i][t] = lambdas[2*i+t];
            energyData[i][t] = sliceEnergies[2*i+t];
        }
    OpenMM::NeighborList list;
    OpenMM::Vec3 box[3];
    for (int r = 0; r < 3; r++)
        box[r] = OpenMM::Vec3(s.box[r][0], s.box[r][1], s.box[r][2]);

    OpenMMLab::ReferenceSlicedLJCoulombIxn clj;
    bool periodic = s.method >= SNB_CUTOFF_PERIODIC;
    if (s.method != SNB_NO_CUTOFF) {
        list.assign(neighbors.begin(), neighbors.end());
        clj.setUseCutoff(s.cutoff, list, s.rfDielectric);
    }
    if (periodic) {
        clj.setPeriodic(box);
        clj.setPeriodicExceptions(s.periodicExceptions);
    }
    int grid[3] = {s.grid[0], s.grid[1], s.grid[2]}, dgrid[3] = {s.dispersionGrid[0], s.dispersionGrid[1], s.dispersionGrid[2]};
    if (s.method == SNB_EWALD)
        clj.setUseEwald(s.alpha, s.kmax[0], s.kmax[1], s.kmax[2]);
    if (s.method == SNB_PME)
        clj.setUsePME(s.alpha, grid);
    if (s.method == SNB_LJPME) {
        clj.setUsePME(s.alpha, grid);
        clj.setUseLJPME(s.dispersionAlpha, dgrid);
    }
    if (s.useSwitch)
        clj.setUseSwitchingFunction(s.switchingDistance);
    clj.calculatePairIxn(numAtoms, posData, numSubsets, subsetData, paramData, lambdaData, exclusions, forceData,
                         energyData, includeDirect, includeReciprocal);
    for (int i = 0; i < numAtoms; i++)
        for (int k = 0; k < 3; k++)
            forces[3*i+k] = forceData[i][k];
    for (int i = 0; i < numSlices; i++)
        for (int t = 0; t < 2; t++)
            sliceEnergies[2*i+t] = energyData[i][t];
}

void backend14Ixn(const PairSettings& s, bool periodic, int i, int j, const double* pos,
                  const double* params, double* forces, const double* lambdas, double* sliceEnergies) {
    // The reference class indexes whole-system vectors; give it a 2-atom system.
    std::vector<int> indices = {0, 1};
    std::vector<OpenMM::Vec3> posData = {OpenMM::Vec3(pos[3*i], pos[3*i+1], pos[3*i+2]), OpenMM::Vec3(pos[3*j], pos[3*j+1], pos[3*j+2])};
    std::vector<OpenMM::Vec3> forceData(2);
    std::vector<double> p(params, params+3), lam(lambdas, lambdas+2), e(sliceEnergies, sliceEnergies+2);
    OpenMMLab::ReferenceSlicedLJCoulomb14 nonbonded14;
    if (periodic) {
        OpenMM::Vec3 box[3];
        for (int r = 0; r < 3; r++)
            box[r] = OpenMM::Vec3(s.box[r][0], s.box[r][1], s.box[r][2]);
        nonbonded14.setPeriodic(box);
    }
    nonbonded14.calculateBondIxn(indices, posData, p, forceData, lam, e);
    for (int k = 0; k < 3; k++) {
        forces[3*i+k] += forceData[0][k];
        forces[3*j+k] += forceData[1][k];
    }
    sliceEnergies[0] = e[0];
    sliceEnergies[1] = e[1];
}

const char* backendName() { return "reference"; }

}  // namespace snb_oracle
