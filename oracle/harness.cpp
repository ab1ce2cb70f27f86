/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the SlicedNonbondedForce path.
 * The product (openmm-lab_b200/) never links, imports or executes this file.
 *
 * harness.cpp restates the glue of the reference's Reference-platform kernel,
 * which cannot be compiled here because it needs OpenMM (absent, un-vendored):
 *
 *   ReferenceCalcSlicedNonbondedForceKernel::initialize
 *       platforms/reference/src/ReferenceOpenMMLabKernels.cpp:65-191
 *   ReferenceCalcSlicedNonbondedForceKernel::execute            :193-261
 *   ReferenceCalcSlicedNonbondedForceKernel::copyParametersToContext :263-312
 *   ReferenceCalcSlicedNonbondedForceKernel::computeParameters  :332-385
 *   SlicedNonbondedForceImpl::calcDispersionCorrections / evalIntegral
 *       openmmapi/src/SlicedNonbondedForceImpl.cpp:150-185,263-354
 *   NonbondedForceImpl::calcPMEParameters / calcEwaldParameters (OpenMM 8.1,
 *       third party; call sites ReferenceOpenMMLabKernels.cpp:166,171,176,178)
 *   computeNeighborListVoxelHash (OpenMM 8.1, third party; call site :203) --
 *       restated as a cell-list search with the same SET semantics: all i<j,
 *       non-excluded, periodic-image distance <= cutoff.
 *
 * This restatement is checked against the real thing: plugin/tests/diff_own_vs_oracle.cpp
 * runs the reference's own ReferenceOpenMMLabKernels.cpp (compiled unchanged on a stand-in
 * OpenMM) and this harness on the same random systems, six methods, offsets, scaling
 * parameters, derivatives, switching, dispersion correction: equal to 4e-14.
 *
 * The arithmetic itself (pair loop, Ewald/PME/LJPME, exclusion corrections,
 * 1-4s) is delegated to a back end (backend.h): the reference's own sources
 * (oracle/_ref build) or the from-scratch port (oracle/port_backend.cpp).
 *
 * Deliberate deviation, stated: the sliced dispersion-correction coefficients
 * are computed in double/64-bit; the reference multiplies in `int`
 * (SlicedNonbondedForceImpl.cpp:346,351) and overflows for N > 16383.  Below
 * that size the two agree bit for bit in exact arithmetic.
 */
#include "slicednb.h"
#include "snb_constants.h"
#include "backend.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

using namespace snb_oracle;

namespace {

thread_local std::string g_lastError;

struct Oracle {
    int numParticles, numSubsets, numSlices, num14, numGlobals;
    int method;
    double cutoff, switchingDistance, rfDielectric, ewaldAlpha, dispersionAlpha;
    bool useSwitch, exceptionsArePeriodic, useDispersionCorrection;
    int kmax[3], grid[3], dgrid[3];
    std::vector<int> subsets;
    std::vector<std::set<int> > exclusions;
    std::vector<int> nb14Exception;                 // exception index of each 1-4
    std::vector<std::array<int, 2> > bonded14Index;
    std::vector<int> bonded14Slice;
    std::vector<std::array<double, 3> > baseParticleParams, baseExceptionParams;
    // offsets: (param, index) -> scale; map semantics as in the reference (last write wins)
    std::map<std::pair<int, int>, std::array<double, 3> > particleOffsets, exceptionOffsets;
    // sliceScaling[slice][term] = global parameter index or -1; hasDerivative likewise
    std::vector<std::array<int, 2> > sliceParam;
    std::vector<std::array<bool, 2> > sliceHasDeriv;
    std::vector<int> derivativeParams;
    std::vector<double> dispersionCoefficients;
    PairList neighbors;                              // of the last execute
    bool neighborsValid;
};

double evalIntegral(double r, double rs, double rc, double sigma) {
    // SlicedNonbondedForceImpl.cpp:150-185 (indefinite integral of LJ x switch)
    double A = 1/(rc-rs), A2 = A*A, A3 = A2*A;
    double sig2 = sigma*sigma, sig6 = sig2*sig2*sig2;
    double rs2 = rs*rs, rs3 = rs*rs2;
    double r2 = r*r, r3 = r*r2, r4 = r*r3, r5 = r*r4, r6 = r*r5, r9 = r3*r6;
    return sig6*A3*((
        sig6*(
            + rs3*28*(6*rs2*A2 + 15*rs*A + 10)
            - r*rs2*945*(rs2*A2 + 2*rs*A + 1)
            + r2*rs*1080*(2*rs2*A2 + 3*rs*A + 1)
            - r3*420*(6*rs2*A2 + 6*rs*A + 1)
            + r4*756*(2*rs*A2 + A)
            - r5*378*A2)
        -r6*(
            + rs3*84*(6*rs2*A2 + 15*rs*A + 10)
            - r*rs2*3780*(rs2*A2 + 2*rs*A + 1)
            + r2*rs*7560*(2*rs2*A2 + 3*rs*A + 1))
        )/(252*r9)
     - log(r)*10*(6*rs2*A2 + 6*rs*A + 1)
     + r*15*(2*rs*A2 + A)
     - r2*3*A2
    );
}

std::vector<double> calcDispersionCorrections(const snb_desc& d) {
    // SlicedNonbondedForceImpl.cpp:263-354, counts in double (see file header)
    int numSlices = d.num_subsets*(d.num_subsets+1)/2;
    std::vector<double> out(numSlices, 0.0);
    if (d.method == SNB_NO_CUTOFF || d.method == SNB_CUTOFF_NON_PERIODIC)
        return out;
    int n = d.num_particles;
    std::vector<double> sigma(d.sigma, d.sigma+n), epsilon(d.epsilon, d.epsilon+n);
    for (int k = 0; k < d.num_particle_offsets; k++) {
        int p = d.particle_offset_index[2*k], idx = d.particle_offset_index[2*k+1];
        double v = d.global_parameter_defaults[p];
        sigma[idx] += v*d.particle_offset_scale[3*k+1];
        epsilon[idx] += v*d.particle_offset_scale[3*k+2];
    }
    typedef std::tuple<double, double, int> ParticleClass;
    std::map<ParticleClass, long long> classCounts;
    for (int i = 0; i < n; i++)
        classCounts[std::make_tuple(sigma[i], epsilon[i], d.subset[i])]++;
    std::vector<double> sum1(numSlices, 0), sum2(numSlices, 0), sum3(numSlices, 0);
    bool useSwitch = d.use_switching_function;
    double cutoff = d.cutoff, switchDist = d.switching_distance;
    for (auto& entry : classCounts) {
        double sig = std::get<0>(entry.first), eps = std::get<1>(entry.first);
        int subset = std::get<2>(entry.first);
        double count = (double) (entry.second*(entry.second+1)/2);
        double sig2 = sig*sig, sig6 = sig2*sig2*sig2;
        int slice = subset*(subset+3)/2;
        sum1[slice] += count*eps*sig6*sig6;
        sum2[slice] += count*eps*sig6;
        if (useSwitch)
            sum3[slice] += count*eps*(evalIntegral(cutoff, switchDist, cutoff, sig)-evalIntegral(switchDist, switchDist, cutoff, sig));
    }
    for (auto c1 = classCounts.begin(); c1 != classCounts.end(); ++c1)
        for (auto c2 = classCounts.begin(); c2 != c1; ++c2) {
            double sig = 0.5*(std::get<0>(c1->first)+std::get<0>(c2->first));
            double eps = sqrt(std::get<1>(c1->first)*std::get<1>(c2->first));
            int slice = snb_slice_index(std::get<2>(c1->first), std::get<2>(c2->first));
            double count = (double) c1->second*(double) c2->second;
            double sig2 = sig*sig, sig6 = sig2*sig2*sig2;
            sum1[slice] += count*eps*sig6*sig6;
            sum2[slice] += count*eps*sig6;
            if (useSwitch)
                sum3[slice] += count*eps*(evalIntegral(cutoff, switchDist, cutoff, sig)-evalIntegral(switchDist, switchDist, cutoff, sig));
        }
    double np = (double) n;
    double numInteractions = (double) (((long long) n*((long long) n+1))/2);
    for (int s = 0; s < numSlices; s++) {
        sum1[s] /= numInteractions; sum2[s] /= numInteractions; sum3[s] /= numInteractions;
        out[s] = 8*np*np*M_PI*(sum1[s]/(9*pow(cutoff, 9))-sum2[s]/(3*pow(cutoff, 3))+sum3[s]);
    }
    return out;
}

// NonbondedForceImpl::calcPMEParameters (OpenMM 8.1): explicit parameters win,
// otherwise alpha = sqrt(-ln(2 tol))/rc, n = max(ceil(2 alpha L/(3 tol^(1/5))), 6)
// (dispersion grid: without the factor 2).
void calcPMEParameters(const snb_desc& d, bool lj, double& alpha, int grid[3]) {
    alpha = lj ? d.dispersion_alpha : d.ewald_alpha;
    const int32_t* g = lj ? d.dispersion_grid : d.pme_grid;
    grid[0] = g[0]; grid[1] = g[1]; grid[2] = g[2];
    if (alpha == 0.0) {
        double tol = d.ewald_error_tolerance;
        alpha = (1.0/d.cutoff)*std::sqrt(-log(2.0*tol));
        for (int k = 0; k < 3; k++) {
            double L = d.default_box[4*k];
            int n = lj ? (int) ceil(alpha*L/(3*pow(tol, 0.2))) : (int) ceil(2*alpha*L/(3*pow(tol, 0.2)));
            grid[k] = std::max(n, 6);
        }
    }
}

// NonbondedForceImpl::calcEwaldParameters (OpenMM 8.1).
struct EwaldErrorFunction {
    double width, alpha, target;
    double operator()(int kmax) const {
        double temp = kmax*M_PI/(width*alpha);
        return target-0.05*sqrt(width*alpha)*kmax*exp(-temp*temp);
    }
};
int findZero(const EwaldErrorFunction& f, int initialGuess) {
    int arg = initialGuess;
    double value = f(arg);
    if (value > 0.0) {
        while (value > 0.0 && arg > 0)
            value = f(--arg);
        return arg+1;
    }
    while (value < 0.0)
        value = f(++arg);
    return arg;
}
void calcEwaldParameters(const snb_desc& d, double& alpha, int kmax[3]) {
    double tol = d.ewald_error_tolerance;
    alpha = (1.0/d.cutoff)*std::sqrt(-log(2.0*tol));
    for (int k = 0; k < 3; k++) {
        EwaldErrorFunction f = {d.default_box[4*k], alpha, tol};
        kmax[k] = findZero(f, 10);
        if (kmax[k]%2 == 0)
            kmax[k]++;
    }
}

void validate(const snb_desc& d) {
    if (d.struct_size != (int32_t) sizeof(snb_desc))
        throw std::invalid_argument("snb_desc.struct_size does not match this library");
    if (d.num_particles < 0 || d.num_subsets < 1)
        throw std::invalid_argument("bad particle or subset count");
    if (d.method < SNB_NO_CUTOFF || d.method > SNB_LJPME)
        throw std::invalid_argument("unknown nonbonded method");
    for (int i = 0; i < d.num_particles; i++)
        if (d.subset[i] < 0 || d.subset[i] >= d.num_subsets)
            throw std::invalid_argument("particle subset out of range");
}

std::vector<int> find14s(const snb_desc& d) {
    std::set<int> withOffsets;
    for (int k = 0; k < d.num_exception_offsets; k++)
        withOffsets.insert(d.exception_offset_index[2*k+1]);
    std::vector<int> nb14s;
    for (int i = 0; i < d.num_exceptions; i++) {
        double qq = d.exception_params[3*i], eps = d.exception_params[3*i+2];
        if (qq != 0.0 || eps != 0.0 || withOffsets.count(i))
            nb14s.push_back(i);
    }
    return nb14s;
}

void load14s(Oracle& o, const snb_desc& d) {
    for (int k = 0; k < o.num14; k++) {
        int e = o.nb14Exception[k];
        int p1 = d.exception_particles[2*e], p2 = d.exception_particles[2*e+1];
        o.baseExceptionParams[k] = {d.exception_params[3*e], d.exception_params[3*e+1], d.exception_params[3*e+2]};
        o.bonded14Index[k] = {p1, p2};
        o.bonded14Slice[k] = snb_slice_index(o.subsets[p1], o.subsets[p2]);
    }
}

Oracle* createOracle(const snb_desc& d) {
    validate(d);
    Oracle* o = new Oracle();
    o->numParticles = d.num_particles;
    o->numSubsets = d.num_subsets;
    o->numSlices = d.num_subsets*(d.num_subsets+1)/2;
    o->numGlobals = d.num_global_parameters;
    o->subsets.assign(d.subset, d.subset+d.num_particles);
    o->sliceParam.assign(o->numSlices, {-1, -1});
    o->sliceHasDeriv.assign(o->numSlices, {false, false});
    o->derivativeParams.assign(d.derivative_parameters, d.derivative_parameters+d.num_derivatives);
    std::set<int> requested(o->derivativeParams.begin(), o->derivativeParams.end());
    for (int k = 0; k < d.num_scaling_parameters; k++) {
        const int32_t* sp = d.scaling_parameters+5*k;
        int slice = snb_slice_index(sp[1], sp[2]);
        bool hasDeriv = requested.count(sp[0]) != 0;
        if (sp[3]) { o->sliceParam[slice][0] = sp[0]; o->sliceHasDeriv[slice][0] = hasDeriv; }
        if (sp[4]) { o->sliceParam[slice][1] = sp[0]; o->sliceHasDeriv[slice][1] = hasDeriv; }
    }
    o->exclusions.resize(d.num_particles);
    for (int i = 0; i < d.num_exceptions; i++) {
        int p1 = d.exception_particles[2*i], p2 = d.exception_particles[2*i+1];
        o->exclusions[p1].insert(p2);
        o->exclusions[p2].insert(p1);
    }
    o->nb14Exception = find14s(d);
    o->num14 = (int) o->nb14Exception.size();
    std::map<int, int> nb14Index;
    for (int k = 0; k < o->num14; k++)
        nb14Index[o->nb14Exception[k]] = k;
    o->bonded14Index.resize(o->num14);
    o->bonded14Slice.resize(o->num14);
    o->baseExceptionParams.resize(o->num14);
    o->baseParticleParams.resize(d.num_particles);
    for (int i = 0; i < d.num_particles; i++)
        o->baseParticleParams[i] = {d.charge[i], d.sigma[i], d.epsilon[i]};
    load14s(*o, d);
    for (int k = 0; k < d.num_particle_offsets; k++)
        o->particleOffsets[std::make_pair(d.particle_offset_index[2*k], d.particle_offset_index[2*k+1])] =
            {d.particle_offset_scale[3*k], d.particle_offset_scale[3*k+1], d.particle_offset_scale[3*k+2]};
    for (int k = 0; k < d.num_exception_offsets; k++)
        o->exceptionOffsets[std::make_pair(d.exception_offset_index[2*k], nb14Index[d.exception_offset_index[2*k+1]])] =
            {d.exception_offset_scale[3*k], d.exception_offset_scale[3*k+1], d.exception_offset_scale[3*k+2]};
    o->method = d.method;
    o->cutoff = d.cutoff;
    o->useSwitch = (d.method != SNB_NO_CUTOFF) && d.use_switching_function;
    o->switchingDistance = d.switching_distance;
    o->ewaldAlpha = o->dispersionAlpha = 0;
    for (int k = 0; k < 3; k++) o->kmax[k] = o->grid[k] = o->dgrid[k] = 0;
    if (d.method == SNB_EWALD)
        calcEwaldParameters(d, o->ewaldAlpha, o->kmax);
    else if (d.method == SNB_PME)
        calcPMEParameters(d, false, o->ewaldAlpha, o->grid);
    else if (d.method == SNB_LJPME) {
        calcPMEParameters(d, false, o->ewaldAlpha, o->grid);
        calcPMEParameters(d, true, o->dispersionAlpha, o->dgrid);
        o->useSwitch = false;
    }
    if (d.method == SNB_NO_CUTOFF || d.method == SNB_CUTOFF_NON_PERIODIC)
        o->exceptionsArePeriodic = false;
    else
        o->exceptionsArePeriodic = d.exceptions_use_pbc != 0;
    o->rfDielectric = d.reaction_field_dielectric;
    o->useDispersionCorrection = d.use_dispersion_correction != 0;
    if (o->useDispersionCorrection)
        o->dispersionCoefficients = calcDispersionCorrections(d);
    else
        o->dispersionCoefficients.assign(o->numSlices, 0.0);
    o->neighborsValid = false;
    return o;
}

void updateOracle(Oracle& o, const snb_desc& d) {
    // ReferenceOpenMMLabKernels.cpp:263-312
    if (d.num_particles != o.numParticles)
        throw std::length_error("updateParametersInContext: The number of particles has changed");
    validate(d);
    o.subsets.assign(d.subset, d.subset+d.num_particles);
    std::vector<int> nb14s = find14s(d);
    if ((int) nb14s.size() != o.num14)
        throw std::domain_error("updateParametersInContext: The number of non-excluded exceptions has changed");
    o.nb14Exception = nb14s;
    for (int i = 0; i < d.num_particles; i++)
        o.baseParticleParams[i] = {d.charge[i], d.sigma[i], d.epsilon[i]};
    load14s(o, d);
    if (d.use_dispersion_correction && (d.method == SNB_CUTOFF_PERIODIC || d.method == SNB_EWALD || d.method == SNB_PME))
        o.dispersionCoefficients = calcDispersionCorrections(d);
}

// ---- neighbour list (set semantics of computeNeighborListVoxelHash) -----------

inline void deltaPeriodic(const double* pi, const double* pj, const double box[3][3], double d[3]) {
    d[0] = pj[0]-pi[0]; d[1] = pj[1]-pi[1]; d[2] = pj[2]-pi[2];
    double s = floor(d[2]/box[2][2]+0.5);
    d[0] -= box[2][0]*s; d[1] -= box[2][1]*s; d[2] -= box[2][2]*s;
    s = floor(d[1]/box[1][1]+0.5);
    d[0] -= box[1][0]*s; d[1] -= box[1][1]*s;
    s = floor(d[0]/box[0][0]+0.5);
    d[0] -= box[0][0]*s;
}

void buildNeighborList(const Oracle& o, const double* pos, const double box[3][3], bool periodic, PairList& out) {
    out.clear();
    int n = o.numParticles;
    double cutoff2 = o.cutoff*o.cutoff;
    int nc[3];
    double lo[3] = {0, 0, 0}, inv[3];
    double recip[3][3] = {{0}};
    if (periodic) {
        // reciprocal (lower triangular) so that frac = pos * recip
        double det = box[0][0]*box[1][1]*box[2][2];
        recip[0][0] = box[1][1]*box[2][2]/det;
        recip[1][0] = -box[1][0]*box[2][2]/det; recip[1][1] = box[0][0]*box[2][2]/det;
        recip[2][0] = (box[1][0]*box[2][1]-box[1][1]*box[2][0])/det; recip[2][1] = -box[0][0]*box[2][1]/det; recip[2][2] = box[0][0]*box[1][1]/det;
        // perpendicular widths of the cell along each lattice direction
        double w[3];
        w[2] = box[2][2];
        w[1] = box[1][1]*box[2][2]/sqrt(box[2][2]*box[2][2]+box[2][1]*box[2][1]);
        double cx = box[1][1]*box[2][2], cy = -box[1][0]*box[2][2], cz = box[1][0]*box[2][1]-box[1][1]*box[2][0];
        w[0] = det/sqrt(cx*cx+cy*cy+cz*cz);
        for (int k = 0; k < 3; k++)
            nc[k] = std::max(1, std::min(256, (int) floor(w[k]/o.cutoff)));
    }
    else {
        double hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = 1e300; hi[k] = -1e300; }
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], pos[3*i+k]); hi[k] = std::max(hi[k], pos[3*i+k]); }
        for (int k = 0; k < 3; k++) {
            nc[k] = n == 0 ? 1 : std::max(1, std::min(256, (int) floor((hi[k]-lo[k])/o.cutoff)));
            inv[k] = (n == 0 || hi[k] == lo[k]) ? 0.0 : nc[k]/(hi[k]-lo[k]);
        }
    }
    int ncells = nc[0]*nc[1]*nc[2];
    std::vector<int> cellOf(n), cellStart(ncells+1, 0), order(n);
    for (int i = 0; i < n; i++) {
        int c[3];
        for (int k = 0; k < 3; k++) {
            if (periodic) {
                double f = pos[3*i]*recip[0][k]+pos[3*i+1]*recip[1][k]+pos[3*i+2]*recip[2][k];
                f -= floor(f);
                c[k] = std::min(nc[k]-1, (int) (f*nc[k]));
            }
            else
                c[k] = std::max(0, std::min(nc[k]-1, (int) ((pos[3*i+k]-lo[k])*inv[k])));
        }
        cellOf[i] = (c[0]*nc[1]+c[1])*nc[2]+c[2];
        cellStart[cellOf[i]+1]++;
    }
    for (int c = 0; c < ncells; c++)
        cellStart[c+1] += cellStart[c];
    std::vector<int> fill(cellStart.begin(), cellStart.end()-1);
    for (int i = 0; i < n; i++)
        order[fill[cellOf[i]]++] = i;
    for (int cxi = 0; cxi < nc[0]; cxi++)
    for (int cyi = 0; cyi < nc[1]; cyi++)
    for (int czi = 0; czi < nc[2]; czi++) {
        int c = (cxi*nc[1]+cyi)*nc[2]+czi;
        std::set<int> nbrCells;
        for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
            int x = cxi+dx, y = cyi+dy, z = czi+dz;
            if (periodic) {
                x = (x+nc[0])%nc[0]; y = (y+nc[1])%nc[1]; z = (z+nc[2])%nc[2];
            }
            else if (x < 0 || y < 0 || z < 0 || x >= nc[0] || y >= nc[1] || z >= nc[2])
                continue;
            nbrCells.insert((x*nc[1]+y)*nc[2]+z);
        }
        for (int c2 : nbrCells) {
            if (c2 < c)
                continue;
            for (int a = cellStart[c]; a < cellStart[c+1]; a++) {
                int i = order[a];
                for (int b = (c2 == c ? a+1 : cellStart[c2]); b < cellStart[c2+1]; b++) {
                    int j = order[b];
                    // the displacement is always taken from the lower to the higher index, so that the
                    // predicate is a function of the unordered pair (the CUDA path does the same)
                    int lo_ = std::min(i, j), hi_ = std::max(i, j);
                    double d[3];
                    if (periodic)
                        deltaPeriodic(pos+3*lo_, pos+3*hi_, box, d);
                    else {
                        d[0] = pos[3*hi_]-pos[3*lo_]; d[1] = pos[3*hi_+1]-pos[3*lo_+1]; d[2] = pos[3*hi_+2]-pos[3*lo_+2];
                    }
                    double r2 = d[0]*d[0]+d[1]*d[1]+d[2]*d[2];
                    if (r2 > cutoff2)
                        continue;
                    if (!o.exclusions[lo_].empty() && o.exclusions[lo_].count(hi_))
                        continue;
                    out.push_back(std::make_pair(lo_, hi_));
                }
            }
        }
    }
}

void bruteNeighborList(const Oracle& o, const double* pos, const double box[3][3], bool periodic, PairList& out) {
    out.clear();
    double cutoff2 = o.cutoff*o.cutoff;
    for (int i = 0; i < o.numParticles; i++)
        for (int j = i+1; j < o.numParticles; j++) {
            double d[3];
            if (periodic)
                deltaPeriodic(pos+3*i, pos+3*j, box, d);
            else {
                d[0] = pos[3*j]-pos[3*i]; d[1] = pos[3*j+1]-pos[3*i+1]; d[2] = pos[3*j+2]-pos[3*i+2];
            }
            if (d[0]*d[0]+d[1]*d[1]+d[2]*d[2] > cutoff2 || o.exclusions[i].count(j))
                continue;
            out.push_back(std::make_pair(i, j));
        }
}

void loadBox(const double* box, double out[3][3]) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            out[r][c] = box ? box[3*r+c] : 0.0;
}

double executeOracle(Oracle& o, const double* pos, const double* boxIn, const double* globals, int flags,
                     double* forcesOut, double* sliceEnergiesOut, double* derivs) {
    bool includeForces = flags & SNB_INCLUDE_FORCES, includeEnergy = flags & SNB_INCLUDE_ENERGY;
    bool includeDirect = flags & SNB_INCLUDE_DIRECT, includeReciprocal = flags & SNB_INCLUDE_RECIPROCAL;
    int n = o.numParticles, L = o.numSlices;

    // computeParameters (:332-385)
    std::vector<double> lambdas(2*L);
    for (int s = 0; s < L; s++)
        for (int t = 0; t < 2; t++)
            lambdas[2*s+t] = o.sliceParam[s][t] < 0 ? 1.0 : globals[o.sliceParam[s][t]];
    std::vector<double> charges(n), sigmas(n), epsilons(n);
    for (int i = 0; i < n; i++) {
        charges[i] = o.baseParticleParams[i][0]; sigmas[i] = o.baseParticleParams[i][1]; epsilons[i] = o.baseParticleParams[i][2];
    }
    for (auto& off : o.particleOffsets) {
        double v = globals[off.first.first];
        int idx = off.first.second;
        charges[idx] += v*off.second[0]; sigmas[idx] += v*off.second[1]; epsilons[idx] += v*off.second[2];
    }
    std::vector<double> particleParams(3*(size_t) n);
    for (int i = 0; i < n; i++) {
        particleParams[3*i] = 0.5*sigmas[i];
        particleParams[3*i+1] = 2.0*sqrt(epsilons[i]);
        particleParams[3*i+2] = charges[i];
    }
    std::vector<double> q14(o.num14), s14(o.num14), e14(o.num14);
    for (int k = 0; k < o.num14; k++) {
        q14[k] = o.baseExceptionParams[k][0]; s14[k] = o.baseExceptionParams[k][1]; e14[k] = o.baseExceptionParams[k][2];
    }
    for (auto& off : o.exceptionOffsets) {
        double v = globals[off.first.first];
        int idx = off.first.second;
        q14[idx] += v*off.second[0]; s14[idx] += v*off.second[1]; e14[idx] += v*off.second[2];
    }

    // execute (:193-261)
    PairSettings s;
    s.method = o.method; s.cutoff = o.cutoff; s.useSwitch = o.useSwitch; s.switchingDistance = o.switchingDistance;
    s.rfDielectric = o.rfDielectric; s.periodicExceptions = o.exceptionsArePeriodic;
    s.alpha = o.ewaldAlpha; s.dispersionAlpha = o.dispersionAlpha;
    for (int k = 0; k < 3; k++) { s.grid[k] = o.grid[k]; s.dispersionGrid[k] = o.dgrid[k]; s.kmax[k] = o.kmax[k]; }
    loadBox(boxIn, s.box);
    bool isPeriodic = o.method >= SNB_CUTOFF_PERIODIC;
    if (o.method != SNB_NO_CUTOFF) {
        buildNeighborList(o, pos, s.box, isPeriodic, o.neighbors);
        o.neighborsValid = true;
    }
    if (isPeriodic) {
        double minAllowedSize = 1.999999*o.cutoff;
        if (s.box[0][0] < minAllowedSize || s.box[1][1] < minAllowedSize || s.box[2][2] < minAllowedSize)
            throw std::range_error("The periodic box size has decreased to less than twice the nonbonded cutoff.");
    }
    std::vector<double> scratch;
    double* forces = forcesOut;
    if (!includeForces || forces == NULL) {
        scratch.assign(3*(size_t) n, 0.0);
        forces = scratch.data();
    }
    std::vector<double> sliceEnergies(2*L, 0.0);
    backendPairIxn(s, n, pos, o.numSubsets, o.subsets.data(), particleParams.data(), lambdas.data(), o.exclusions,
                   o.neighbors, forces, sliceEnergies.data(), includeDirect, includeReciprocal);
    if (includeDirect) {
        for (int k = 0; k < o.num14; k++) {
            int slice = o.bonded14Slice[k];
            double p[3] = {s14[k], 4.0*e14[k], q14[k]};
            backend14Ixn(s, o.exceptionsArePeriodic, o.bonded14Index[k][0], o.bonded14Index[k][1], pos, p, forces,
                         &lambdas[2*slice], &sliceEnergies[2*slice]);
        }
        if (o.method == SNB_CUTOFF_PERIODIC || o.method == SNB_EWALD || o.method == SNB_PME) {
            double volume = s.box[0][0]*s.box[1][1]*s.box[2][2];
            for (int slice = 0; slice < L; slice++)
                sliceEnergies[2*slice+1] += o.dispersionCoefficients[slice]/volume;
        }
    }
    double energy = 0;
    if (includeEnergy)
        for (int i = 0; i < 2*L; i++)
            energy += lambdas[i]*sliceEnergies[i];
    for (int slice = 0; slice < L; slice++)
        for (int t = 0; t < 2; t++)
            if (o.sliceHasDeriv[slice][t])
                for (size_t dIdx = 0; dIdx < o.derivativeParams.size(); dIdx++)
                    if (o.derivativeParams[dIdx] == o.sliceParam[slice][t])
                        derivs[dIdx] += sliceEnergies[2*slice+t];
    if (sliceEnergiesOut)
        std::copy(sliceEnergies.begin(), sliceEnergies.end(), sliceEnergiesOut);
    return energy;
}

int fail(int code, const std::exception& e) {
    g_lastError = e.what();
    return code;
}

int dumpPairs(PairList pairs, int32_t** out, int64_t* count) {
    std::sort(pairs.begin(), pairs.end());
    *count = (int64_t) pairs.size();
    *out = (int32_t*) malloc(std::max<size_t>(1, pairs.size())*2*sizeof(int32_t));
    for (size_t k = 0; k < pairs.size(); k++) {
        (*out)[2*k] = pairs[k].first;
        (*out)[2*k+1] = pairs[k].second;
    }
    return SNB_OK;
}

}  // namespace

#define SNB_TRY try {
#define SNB_CATCH \
    } catch (const std::range_error& e) { return fail(SNB_ERR_BOX_TOO_SMALL, e); } \
      catch (const std::length_error& e) { return fail(SNB_ERR_PARTICLE_COUNT, e); } \
      catch (const std::domain_error& e) { return fail(SNB_ERR_EXCEPTION_SET, e); } \
      catch (const std::invalid_argument& e) { return fail(SNB_ERR_INVALID, e); } \
      catch (const std::exception& e) { return fail(SNB_ERR_INVALID, e); }

extern "C" {

int snb_oracle_create(const snb_desc* desc, void** out) {
    SNB_TRY
    *out = createOracle(*desc);
    return SNB_OK;
    SNB_CATCH
}

void snb_oracle_destroy(void* h) { delete (Oracle*) h; }

int snb_oracle_update_parameters(void* h, const snb_desc* desc) {
    SNB_TRY
    updateOracle(*(Oracle*) h, *desc);
    return SNB_OK;
    SNB_CATCH
}

int snb_oracle_execute(void* h, const double* positions, const double* box, const double* globals, int flags,
                       double* forces, double* energy, double* slice_energies, double* derivatives) {
    SNB_TRY
    double e = executeOracle(*(Oracle*) h, positions, box, globals, flags, forces, slice_energies, derivatives);
    if (energy) *energy = e;
    return SNB_OK;
    SNB_CATCH
}

int snb_oracle_get_pme_parameters(void* h, double* alpha, int* nx, int* ny, int* nz) {
    Oracle& o = *(Oracle*) h;
    if (o.method != SNB_PME && o.method != SNB_LJPME) {
        g_lastError = "getPMEParametersInContext: This Context is not using PME or LJPME";
        return SNB_ERR_WRONG_METHOD;
    }
    *alpha = o.ewaldAlpha; *nx = o.grid[0]; *ny = o.grid[1]; *nz = o.grid[2];
    return SNB_OK;
}

int snb_oracle_get_ljpme_parameters(void* h, double* alpha, int* nx, int* ny, int* nz) {
    Oracle& o = *(Oracle*) h;
    if (o.method != SNB_LJPME) {
        g_lastError = "getPMEParametersInContext: This Context is not using LJPME";
        return SNB_ERR_WRONG_METHOD;
    }
    *alpha = o.dispersionAlpha; *nx = o.dgrid[0]; *ny = o.dgrid[1]; *nz = o.dgrid[2];
    return SNB_OK;
}

int snb_oracle_get_ewald_parameters(void* h, double* alpha, int* kx, int* ky, int* kz) {
    Oracle& o = *(Oracle*) h;
    *alpha = o.ewaldAlpha; *kx = o.kmax[0]; *ky = o.kmax[1]; *kz = o.kmax[2];
    return SNB_OK;
}

int snb_oracle_dispersion_coefficients(void* h, double* out) {
    Oracle& o = *(Oracle*) h;
    std::copy(o.dispersionCoefficients.begin(), o.dispersionCoefficients.end(), out);
    return SNB_OK;
}

/* sorted (i<j) neighbour list of the last execute */
int snb_oracle_dump_pairs(void* h, int32_t** pairs, int64_t* count) {
    Oracle& o = *(Oracle*) h;
    if (!o.neighborsValid) {
        g_lastError = "no neighbour list has been built";
        return SNB_ERR_INVALID;
    }
    return dumpPairs(o.neighbors, pairs, count);
}

/* number of pairs in the neighbour list of the last execute (full-size checks: no copy, no sort) */
long long snb_oracle_count_pairs(void* h) {
    Oracle& o = *(Oracle*) h;
    return o.neighborsValid ? (long long) o.neighbors.size() : -1;
}

/* O(N^2) enumeration, to validate the cell-list search itself */
int snb_oracle_brute_pairs(void* h, const double* positions, const double* box, int32_t** pairs, int64_t* count) {
    Oracle& o = *(Oracle*) h;
    double b[3][3];
    loadBox(box, b);
    PairList list;
    bruteNeighborList(o, positions, b, o.method >= SNB_CUTOFF_PERIODIC, list);
    return dumpPairs(list, pairs, count);
}

/* smallest |r - cutoff| over all non-excluded pairs near the cutoff shell (guard band check) */
double snb_oracle_cutoff_margin(void* h, const double* positions, const double* box) {
    Oracle& o = *(Oracle*) h;
    double b[3][3];
    loadBox(box, b);
    Oracle wide = o;
    wide.cutoff = o.cutoff*1.02;
    PairList list;
    buildNeighborList(wide, positions, b, o.method >= SNB_CUTOFF_PERIODIC, list);
    double best = 1e300;
    for (auto& p : list) {
        double d[3];
        if (o.method >= SNB_CUTOFF_PERIODIC)
            deltaPeriodic(positions+3*p.first, positions+3*p.second, b, d);
        else
            for (int k = 0; k < 3; k++) d[k] = positions[3*p.second+k]-positions[3*p.first+k];
        best = std::min(best, fabs(sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2])-o.cutoff));
    }
    return best;
}

void snb_oracle_free(void* p) { free(p); }
const char* snb_oracle_last_error(void) { return g_lastError.c_str(); }
const char* snb_oracle_backend(void) { return backendName(); }

}  // extern "C"
