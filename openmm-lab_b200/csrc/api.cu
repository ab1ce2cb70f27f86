// C-ABI of the B200 SlicedNonbondedForce core (include/slicednb.h) and the host-side
// orchestration of one evaluation.  Host logic restated from the reference kernel glue:
//   initialize               platforms/reference/src/ReferenceOpenMMLabKernels.cpp:65-191
//   execute                  :193-261        computeParameters :332-385
//   copyParametersToContext  :263-312        get(LJ)PMEParameters :314-330
//   calcDispersionCorrections openmmapi/src/SlicedNonbondedForceImpl.cpp:263-354 (in 64-bit/double;
//                            the reference overflows `int` for N > 16383, :346,351)
// Scheduling differs from the reference's CUDA platform (CudaOpenMMLabKernels.cpp:888-1112) by
// design: one stream-ordered sequence with no host synchronisation until the final read-back.
#include "kernels.cuh"
#include "comm.cuh"

#include <cufft.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_lastError;

struct SnbException : public std::exception {
    int code;
    std::string msg;
    SnbException(int code, const std::string& msg) : code(code), msg(msg) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

#define CUDA_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    throw SnbException(SNB_ERR_CUDA, std::string(#call)+": "+cudaGetErrorString(e_)); } while (0)
#define NCCL_CHECK(api, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    throw SnbException(SNB_ERR_CUDA, std::string(#call)+": "+(api)->GetErrorString(r_)); } while (0)
#define CUFFT_CHECK(call) do { cufftResult r_ = (call); if (r_ != CUFFT_SUCCESS) \
    throw SnbException(SNB_ERR_CUDA, std::string(#call)+": cuFFT error "+std::to_string((int) r_)); } while (0)

// NVTX range over the host-side enqueue of one phase (shows up in Nsight timelines; a no-op without a tool attached)
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owner = true;
    void view(void* ptr, size_t count) {     // a window into another allocation
        release();
        p = (T*) ptr; n = count; owner = false;
    }
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1)*sizeof(T)));
        n = std::max<size_t>(count, 1);
    }
    void release() { if (p && owner) cudaFree(p); p = nullptr; n = 0; owner = true; }
    void upload(const T* src, size_t count, cudaStream_t s) {
        alloc(count);
        if (count) CUDA_CHECK(cudaMemcpyAsync(p, src, count*sizeof(T), cudaMemcpyHostToDevice, s));
    }
    ~DevBuf() { release(); }
};
template <typename T> struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        if (count <= n && p) return;
        if (p) cudaFreeHost(p);
        CUDA_CHECK(cudaMallocHost(&p, std::max<size_t>(count, 1)*sizeof(T)));
        n = std::max<size_t>(count, 1);
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

double evalIntegral(double r, double rs, double rc, double sigma) {
    // SlicedNonbondedForceImpl.cpp:150-185
    double A = 1/(rc-rs), A2 = A*A, A3 = A2*A;
    double sig2 = sigma*sigma, sig6 = sig2*sig2*sig2;
    double rs2 = rs*rs, rs3 = rs*rs2;
    double r2 = r*r, r3 = r*r2, r4 = r*r3, r5 = r*r4, r6 = r*r5, r9 = r3*r6;
    return sig6*A3*((
        sig6*(rs3*28*(6*rs2*A2+15*rs*A+10)-r*rs2*945*(rs2*A2+2*rs*A+1)+r2*rs*1080*(2*rs2*A2+3*rs*A+1)
              -r3*420*(6*rs2*A2+6*rs*A+1)+r4*756*(2*rs*A2+A)-r5*378*A2)
        -r6*(rs3*84*(6*rs2*A2+15*rs*A+10)-r*rs2*3780*(rs2*A2+2*rs*A+1)+r2*rs*7560*(2*rs2*A2+3*rs*A+1))
        )/(252*r9)-log(r)*10*(6*rs2*A2+6*rs*A+1)+r*15*(2*rs*A2+A)-r2*3*A2);
}

uint64_t hilbertIndex(uint32_t x, uint32_t y, uint32_t z, int bits) {
    // Skilling's transpose form of the 3-D Hilbert curve
    uint32_t X[3] = {x, y, z};
    uint32_t M = 1u<<(bits-1), t;
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        uint32_t P = Q-1;
        for (int i = 0; i < 3; i++) {
            if (X[i]&Q) X[0] ^= P;
            else { t = (X[0]^X[i])&P; X[0] ^= t; X[i] ^= t; }
        }
    }
    for (int i = 1; i < 3; i++) X[i] ^= X[i-1];
    t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2]&Q) t ^= Q-1;
    for (int i = 0; i < 3; i++) X[i] ^= t;
    uint64_t h = 0;
    for (int b = bits-1; b >= 0; b--)
        for (int i = 0; i < 3; i++)
            h = (h<<1)|((X[i]>>b)&1u);
    return h;
}

std::vector<double> bsplineModuli(int n) {
    // ReferencePME.cpp:88-183: |DFT of the integer-sampled order-5 spline|^2 with the <1e-7 fix-up
    const int order = SNB_PME_ORDER;
    double data[SNB_PME_ORDER];
    data[order-1] = 0; data[1] = 0; data[0] = 1;
    for (int k = 3; k < order; k++) {
        double div = 1.0/(k-1.0);
        data[k-1] = 0;
        for (int l = 1; l < k-1; l++)
            data[k-l-1] = div*(l*data[k-l-2]+(k-l)*data[k-l-1]);
        data[0] = div*data[0];
    }
    double div = 1.0/(order-1);
    data[order-1] = 0;
    for (int l = 1; l < order-1; l++)
        data[order-l-1] = div*(l*data[order-l-2]+(order-l)*data[order-l-1]);
    data[0] = div*data[0];
    std::vector<double> bs(std::max(n, order+1), 0.0), mod(n);
    for (int i = 1; i <= order; i++) bs[i] = data[i-1];
    for (int i = 0; i < n; i++) {
        double sc = 0, ss = 0;
        for (int j = 0; j < n; j++) {
            double arg = (2.0*M_PI*i*j)/n;
            sc += bs[j]*cos(arg);
            ss += bs[j]*sin(arg);
        }
        mod[i] = sc*sc+ss*ss;
    }
    for (int i = 0; i < n; i++)
        if (mod[i] < 1.0e-7)
            mod[i] = (mod[(i-1+n)%n]+mod[(i+1)%n])/2;
    return mod;
}

struct PmeGrid {
    int n[3] = {0, 0, 0};
    double alpha = 0;
    cufftHandle fwd = 0, inv = 0;
    bool planned = false;
    DevBuf<char> real, cplx, eterm;     // float or double, by the handle's precision
    DevBuf<long long> fixed;            // deterministic spreading
    DevBuf<float> pad;                  // brick path: padded copy of the grids addressed by the TMA bricks
    alignas(64) CUtensorMap brickMap;
    int nzp = 0;
    bool brick = false;
    DevBuf<double> moduli[3];
    double etermBox[6] = {0, 0, 0, 0, 0, 0};
    bool etermValid = false;
    ~PmeGrid() { if (planned) { cufftDestroy(fwd); cufftDestroy(inv); } }
};

}  // namespace

// Everything that shapes the launches of one evaluation.  Two evaluations with equal keys enqueue the
// same nodes with the same arguments (the box and the lambdas travel through StepParams in device
// memory), so the second one onwards replays a captured CUDA graph.
struct GraphKey {
    int flags, hostPositions, pbcMode, cellDim[3], numCells, needEnergy, broadcast, fusedSort, beginKernel, compactSort;
    size_t maxChunks;
    const void *posqDev, *forcesFixedOut, *jList, *hostPos;
    const void *posqCorrection, *atomIndex;     // snb_execute_device_ordered: the caller's atom order
    long long outStride, positionsDouble, pmeEarly;
    double share, listSkin;
};

#define SNB_ENERGY_FLAGS 32
#define SNB_ENERGY_DOUBLES (32*2*SNB_MAX_SLICES+SNB_ENERGY_FLAGS)   /* 32 replicas of [slice][term] + flags: [0] list overflow,
                                                                      [1] rank 0's reciprocal-space ms, [2+r] rank r's list+direct ms,
                                                                      [10+r] rank r's ms until its forces were ready (r < 8) */

struct snb_handle {
    // ---- snapshot of the force (ReferenceOpenMMLabKernels.cpp:65-191) ----
    int N = 0, NP = 0, NB = 0, S = 1, L = 1, G = 0, method = 0, device = 0;
    double cutoff = 1, switchDist = 0, rfDielectric = 78.3, alpha = 0, dalpha = 0;
    bool useSwitch = false, exceptionsPeriodic = false, useDispersion = false;
    bool doublePrecision = false, deterministic = false;
    int kmax[3] = {0, 0, 0}, numK = 0;  // classic Ewald: reciprocal vectors per axis, size of the half space
    DevBuf<int3> ewaldK;
    DevBuf<double2> ewaldS;
    int directKernelFlags = 0;          // SNB_FLAG_DIRECT_PACKED / _GENERAL, 0 = by system size
    int grid[3] = {0, 0, 0}, dgrid[3] = {0, 0, 0};
    std::vector<int> subsets;
    std::vector<std::array<double, 3>> baseParticle;          // q, sigma, eps
    int X = 0, num14 = 0;
    std::vector<int2> exceptionAtoms;
    std::vector<std::array<double, 3>> baseException;         // chargeProd, sigma, eps
    std::vector<int> is14;
    std::vector<std::array<int, 2>> particleOffsetIdx, exceptionOffsetIdx;   // (param, index)
    std::vector<std::array<double, 3>> particleOffsetScale, exceptionOffsetScale;
    std::vector<std::array<int, 2>> sliceParam;                // [L] (Coul, vdW) global index or -1
    std::vector<std::array<bool, 2>> sliceHasDeriv;
    std::vector<int> derivativeParams;
    std::vector<double> dispersionCoef;
    std::set<int> offsetParams;
    // ---- derived host state ----
    bool paramsDirty = true;
    std::vector<double> lastOffsetValues;
    // computeParameters runs on the device (params.cu): base parameters and the offset CSRs live in HBM
    DevBuf<double> baseParticleD, baseExceptionD, globalsD, selfD, selfPartial;
    DevBuf<int> particleOffsetStart, exceptionOffsetStart;
    DevBuf<double4> particleOffsetsD, exceptionOffsetsD;
    DevBuf<unsigned> paramTicket;
    PinnedBuf<double> hGlobals, hSelf;                         // hSelf: [subset][Coulomb, dispersion] self energies
    double lambdaC[SNB_MAX_SLICES], lambdaV[SNB_MAX_SLICES];
    int cellDim[3] = {0, 0, 0}, numCells = 0;
    int pbcModeOverride = -1;
    // ---- device state ----
    cudaStream_t stream = nullptr, pmeStream = nullptr, bondedStream = nullptr;
    cudaEvent_t evInputs = nullptr, evBondedDone = nullptr;
    DevBuf<StepParams> stepD;
    PinnedBuf<StepParams> stepH;
    bool useGraph = true, graphKeyValid = false;
    GraphKey graphKey;
    cudaGraphExec_t graphExec = nullptr;
    cudaEvent_t evSorted = nullptr, evPmeDone = nullptr, evBegin = nullptr, evEnd = nullptr;
    float lastDeviceMs = 0;
    int numSMs = 148;
    int fusedSort = 1;          // the sort as one cooperative kernel where it applies (SNB_FUSED_SORT=0: separate kernels); cleared if its launch fails
    bool fusedSortUsed = false; // the last enqueueStep launched it
    bool fusedSortFailed = false;
    int pmeFoldedLaunches = 0;  // reciprocal-space terms of the last enqueueStep whose interpolation did the lambda combination itself
    int compactSort = 1;        // sort in three kernels instead of five where it applies (SNB_COMPACT_SORT=0: five)
    int beginKernel = 1;        // k_begin instead of copy + memset + conversion kernel (SNB_BEGIN_KERNEL=0: those three)
    bool beginKernelUsed = false;
    const int* stepHDev = nullptr;
    int hostResults = 1;        // k_finish stores energies / counters into the pinned mirrors (SNB_HOST_RESULTS=0: copies)
    double* hEnergyDev = nullptr;
    int* hCountersDev = nullptr;
    DevBuf<double> posd, paramsD, exception14, extent, energyBuf, forcesOutD;
    DevBuf<float4> params4, posqLocal, sparams, blockExt, posqIn, octetCenter, octetExt, blockCenter;
    DevBuf<double4> posqLocalD, sparamsD;
    DevBuf<int> slotCell;
    DevBuf<int> subsetD, cellRank, cellCount, cellStart, atomCell, atomSlot, sortedAtom, sortedPos,
                exclStart, exclList, is14D, counters, cellScratch, tileSum;
    DevBuf<int2> exceptionAtomsD, jList, cellRange;
    DevBuf<int> chunkOctet;
    DevBuf<long long> forceSorted, forceOrig;
    DevBuf<char> zeroBlock;
    size_t zeroBytes = 0;
    PinnedBuf<double> hPos, hForces, hEnergy;
    PinnedBuf<int> hCounters;
    size_t maxChunks = 0;
    // ---- tile-list reuse (platform property ListSkin, snb_set_list_skin; 0 = rebuilt every evaluation) ----
    double listSkin = 0;                // as requested
    double skinNow = 0;                 // in effect for the evaluation being prepared (a small box clips it: cutoff + skin <= half the box)
    DevBuf<double> posRef;              // positions at the last rebuild
    DevBuf<float4> wrapImage;           // periodic image every atom was wrapped by at the last rebuild
    DevBuf<int> listState;              // [0] chunks of the kept list
    bool listKept = false;              // a list built with this skin, box, capacity and partition is on the device
    double listBox[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, listShare = 0, listBuiltSkin = 0;
    size_t listMaxChunks = 0;
    long long listRebuilds = 0, listEvals = 0;
    PmeGrid pme, dpme;
    // ---- multi-GPU (comm.cuh): tile list and exceptions partitioned, reciprocal space on rank 0 ----
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double rank0Share = -1;             // < 0: cost-model default
    DevBuf<long long> reduceSend, reduceRecv;
    DevBuf<long long> forcePme;         // rank 0: reciprocal-space forces, kept apart from what is reduced ([3][NP], original order)
    DevBuf<double> energyPme;           // rank 0: reciprocal-space slice energies, likewise
    PinnedBuf<double> hEnergyPme;
    PinnedBuf<long long> hEnergyFixed;  // the all-reduced energy table as it arrives (2^32 fixed point)
    cudaEvent_t evPacked = nullptr;     // this rank's contribution is packed (timing; load balance)
    // rank 0 with (next to) no share of the tile list (4 ranks and more): reciprocal space need not wait for this evaluation's sort --
    // the sorted order only gives its kernels locality, any permutation is correct -- so it starts with the inputs, on a COPY of
    // the previous evaluation's order (the sort kernels rewrite the live one meanwhile)
    DevBuf<int> sortedAtomPme;
    bool sortedPmeValid = false, pmeEarly = false;
    bool broadcastPositions = true;     // false: the caller guarantees identical positions on every rank
    bool balanceFrozen = false;         // after a few measured evaluations the partition is fixed (and graphs take over)
    int commEvals = 0;
    cudaEvent_t evDirBegin = nullptr, evDirEnd = nullptr, evRecBegin = nullptr, evRecEnd = nullptr, evReady = nullptr;
    PinnedBuf<double> loadH;            // this rank's measured times of the previous evaluation, sent with the all-reduce
    bool loadValid = false, autoBalance = true;
    double tunedShare = -1;
    // ---- snb_execute_device_ordered: the caller's layout for the evaluation in progress (all null / 0 otherwise) ----
    const void* ordCorrection = nullptr;
    const int* ordAtomIndex = nullptr;
    int ordStride = 0, ordDouble = 0;
    // ---- results of the last execute ----
    std::vector<double> sliceEnergies;
    double lastEnergy = 0;
    std::vector<double> lastGlobals;
    int lastFlags = 0;
    bool listValid = false;
    DirectArgs lastDirect;
    bool timing = false;
    cudaEvent_t ev[SNB_NUM_PHASES+2];
    bool evCreated = false;
    float phaseMs[SNB_NUM_PHASES];
    long long countersOut[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool packedLaunched = false;        // which direct-space kernel the last enqueue chose
    long long paramLaunches = 0;        // k_computeParameters launches so far
};

namespace {

void validateDesc(const snb_desc& d) {
    if (d.struct_size != (int32_t) sizeof(snb_desc))
        throw SnbException(SNB_ERR_INVALID, "snb_desc.struct_size does not match this library");
    if (d.num_particles < 0 || d.num_subsets < 1)
        throw SnbException(SNB_ERR_INVALID, "bad particle or subset count");
    if (d.num_subsets > SNB_MAX_SUBSETS)
        throw SnbException(SNB_ERR_UNSUPPORTED, "more than 8 subsets are not supported by this build");
    if (d.method < SNB_NO_CUTOFF || d.method > SNB_LJPME)
        throw SnbException(SNB_ERR_INVALID, "unknown nonbonded method");
    if (d.method == SNB_EWALD && (d.default_box[3] != 0.0 || d.default_box[6] != 0.0 || d.default_box[7] != 0.0))
        throw SnbException(SNB_ERR_INVALID, "SlicedNonbondedForce: Ewald is not supported with non-rectangular boxes.  Use PME instead.");
    for (int i = 0; i < d.num_particles; i++)
        if (d.subset[i] < 0 || d.subset[i] >= d.num_subsets)
            throw SnbException(SNB_ERR_INVALID, "particle subset out of range");
    if (d.num_exceptions < 0 || d.num_global_parameters < 0 || d.num_particle_offsets < 0 || d.num_exception_offsets < 0 ||
        d.num_scaling_parameters < 0 || d.num_derivatives < 0)
        throw SnbException(SNB_ERR_INVALID, "negative count in the descriptor");
    // every index the host and device loops will use
    for (int k = 0; k < d.num_particle_offsets; k++)
        if (d.particle_offset_index[2*k] < 0 || d.particle_offset_index[2*k] >= d.num_global_parameters ||
            d.particle_offset_index[2*k+1] < 0 || d.particle_offset_index[2*k+1] >= d.num_particles)
            throw SnbException(SNB_ERR_INVALID, "particle parameter offset: parameter or particle index out of range");
    for (int k = 0; k < d.num_exception_offsets; k++)
        if (d.exception_offset_index[2*k] < 0 || d.exception_offset_index[2*k] >= d.num_global_parameters ||
            d.exception_offset_index[2*k+1] < 0 || d.exception_offset_index[2*k+1] >= d.num_exceptions)
            throw SnbException(SNB_ERR_INVALID, "exception parameter offset: parameter or exception index out of range");
    for (int k = 0; k < d.num_derivatives; k++)
        if (d.derivative_parameters[k] < 0 || d.derivative_parameters[k] >= d.num_global_parameters)
            throw SnbException(SNB_ERR_INVALID, "derivative parameter index out of range");
    // user-supplied PME grids: the spreading / interpolation stencils wrap once, which needs at least ORDER points per axis
    if (d.method == SNB_PME || d.method == SNB_LJPME) {
        if (d.ewald_alpha != 0.0 && (d.pme_grid[0] < SNB_PME_ORDER || d.pme_grid[1] < SNB_PME_ORDER || d.pme_grid[2] < SNB_PME_ORDER))
            throw SnbException(SNB_ERR_INVALID, "PME grid dimensions must be at least the spline order (5)");
        if (d.method == SNB_LJPME && d.dispersion_alpha != 0.0 &&
            (d.dispersion_grid[0] < SNB_PME_ORDER || d.dispersion_grid[1] < SNB_PME_ORDER || d.dispersion_grid[2] < SNB_PME_ORDER))
            throw SnbException(SNB_ERR_INVALID, "LJPME grid dimensions must be at least the spline order (5)");
    }
}

std::vector<int> find14s(const snb_desc& d) {
    std::set<int> withOffsets;
    for (int k = 0; k < d.num_exception_offsets; k++)
        withOffsets.insert(d.exception_offset_index[2*k+1]);
    std::vector<int> is14(d.num_exceptions, 0);
    for (int i = 0; i < d.num_exceptions; i++) {
        double qq = d.exception_params[3*i], eps = d.exception_params[3*i+2];
        if (qq != 0.0 || eps != 0.0 || withOffsets.count(i))
            is14[i] = 1;
    }
    return is14;
}

std::vector<double> calcDispersionCorrections(const snb_desc& d) {
    int L = d.num_subsets*(d.num_subsets+1)/2;
    std::vector<double> out(L, 0.0);
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
    std::vector<double> sum1(L, 0), sum2(L, 0), sum3(L, 0);
    bool useSwitch = d.use_switching_function;
    double rc = d.cutoff, rs = d.switching_distance;
    auto add = [&](int slice, double count, double sig, double eps) {
        double s2 = sig*sig, s6 = s2*s2*s2;
        sum1[slice] += count*eps*s6*s6;
        sum2[slice] += count*eps*s6;
        if (useSwitch)
            sum3[slice] += count*eps*(evalIntegral(rc, rs, rc, sig)-evalIntegral(rs, rs, rc, sig));
    };
    for (auto& c : classCounts) {
        int sub = std::get<2>(c.first);
        add(sub*(sub+3)/2, (double) (c.second*(c.second+1)/2), std::get<0>(c.first), std::get<1>(c.first));
    }
    for (auto c1 = classCounts.begin(); c1 != classCounts.end(); ++c1)
        for (auto c2 = classCounts.begin(); c2 != c1; ++c2)
            add(snb_slice_index(std::get<2>(c1->first), std::get<2>(c2->first)), (double) c1->second*(double) c2->second,
                0.5*(std::get<0>(c1->first)+std::get<0>(c2->first)), sqrt(std::get<1>(c1->first)*std::get<1>(c2->first)));
    double np = n, numInteractions = (double) (((long long) n*((long long) n+1))/2);
    for (int s = 0; s < L; s++)
        out[s] = 8*np*np*M_PI*(sum1[s]/numInteractions/(9*pow(rc, 9))-sum2[s]/numInteractions/(3*pow(rc, 3))+sum3[s]/numInteractions);
    return out;
}

void calcPMEParameters(const snb_desc& d, bool lj, double& alpha, int grid[3]) {
    // OpenMM NonbondedForceImpl::calcPMEParameters (third party; call sites ReferenceOpenMMLabKernels.cpp:171,176,178)
    alpha = lj ? d.dispersion_alpha : d.ewald_alpha;
    const int32_t* g = lj ? d.dispersion_grid : d.pme_grid;
    grid[0] = g[0]; grid[1] = g[1]; grid[2] = g[2];
    if (alpha == 0.0) {
        double tol = d.ewald_error_tolerance;
        alpha = (1.0/d.cutoff)*std::sqrt(-log(2.0*tol));
        for (int k = 0; k < 3; k++) {
            double len = d.default_box[4*k];
            int n = lj ? (int) ceil(alpha*len/(3*pow(tol, 0.2))) : (int) ceil(2*alpha*len/(3*pow(tol, 0.2)));
            grid[k] = std::max(n, 6);
        }
    }
}

// OpenMM NonbondedForceImpl::calcEwaldParameters (third party, OpenMM 8.1; call site ReferenceOpenMMLabKernels.cpp:169-173):
// alpha from the tolerance, kmax per axis = smallest odd-rounded k with 0.05 sqrt(L alpha) k exp(-(pi k/(L alpha))^2) < tol
void calcEwaldParameters(const snb_desc& d, double& alpha, int kmax[3]) {
    const double tol = d.ewald_error_tolerance;
    alpha = (1.0/d.cutoff)*std::sqrt(-log(2.0*tol));
    for (int k = 0; k < 3; k++) {
        const double width = d.default_box[4*k];
        auto err = [&](int km) {
            const double t = km*M_PI/(width*alpha);
            return tol-0.05*sqrt(width*alpha)*km*exp(-t*t);
        };
        int arg = 10;
        double value = err(arg);
        if (value > 0.0) {
            while (value > 0.0 && arg > 0)
                value = err(--arg);
            arg++;
        }
        else
            while (value < 0.0)
                value = err(++arg);
        kmax[k] = arg%2 == 0 ? arg+1 : arg;
    }
}

void loadParameters(snb_handle& h, const snb_desc& d) {
    h.subsets.assign(d.subset, d.subset+d.num_particles);
    h.baseParticle.resize(h.N);
    for (int i = 0; i < h.N; i++)
        h.baseParticle[i] = {d.charge[i], d.sigma[i], d.epsilon[i]};
    h.baseException.resize(h.X);
    h.exceptionAtoms.resize(h.X);
    for (int e = 0; e < h.X; e++) {
        h.baseException[e] = {d.exception_params[3*e], d.exception_params[3*e+1], d.exception_params[3*e+2]};
        h.exceptionAtoms[e] = make_int2(d.exception_particles[2*e], d.exception_particles[2*e+1]);
    }
    h.paramsDirty = true;
}

// CSR over `count` items of the offsets given as (parameter, item) pairs; map semantics of the reference (one entry per
// key, a repeated key keeps its last scale; entries of an item ordered by parameter index)
void buildOffsetCsr(int count, const std::vector<std::array<int, 2>>& idx, const std::vector<std::array<double, 3>>& scale,
                    std::vector<int>& start, std::vector<double4>& entries) {
    std::map<std::pair<int, int>, std::array<double, 3>> byKey;     // (item, parameter) -> scale
    for (size_t k = 0; k < idx.size(); k++)
        byKey[{idx[k][1], idx[k][0]}] = scale[k];
    start.assign(count+1, 0);
    entries.clear();
    for (auto& o : byKey)
        start[o.first.first+1]++;
    for (int i = 0; i < count; i++)
        start[i+1] += start[i];
    for (auto& o : byKey)
        entries.push_back(make_double4(o.second[0], o.second[1], o.second[2], (double) o.first.second));
}

// The inputs of the device-side computeParameters (params.cu): base parameters, subsets, offset CSRs.  Called at
// creation and by updateParametersInContext; synchronous (the staging vectors die here).
void uploadBaseParameters(snb_handle& h) {
    std::vector<double> bp(3*(size_t) h.N), be(3*(size_t) h.X);
    for (int i = 0; i < h.N; i++)
        for (int c = 0; c < 3; c++)
            bp[3*(size_t) i+c] = h.baseParticle[i][c];
    for (int e = 0; e < h.X; e++)
        for (int c = 0; c < 3; c++)
            be[3*(size_t) e+c] = h.baseException[e][c];
    h.baseParticleD.upload(bp.data(), bp.size(), h.stream);
    h.baseExceptionD.upload(be.data(), be.size(), h.stream);
    h.subsetD.upload(h.subsets.data(), h.N, h.stream);
    std::vector<int> start;
    std::vector<double4> entries;
    if (!h.particleOffsetIdx.empty()) {
        buildOffsetCsr(h.N, h.particleOffsetIdx, h.particleOffsetScale, start, entries);
        h.particleOffsetStart.upload(start.data(), start.size(), h.stream);
        h.particleOffsetsD.upload(entries.data(), entries.size(), h.stream);
        CUDA_CHECK(cudaStreamSynchronize(h.stream));
    }
    if (!h.exceptionOffsetIdx.empty()) {
        buildOffsetCsr(h.X, h.exceptionOffsetIdx, h.exceptionOffsetScale, start, entries);
        h.exceptionOffsetStart.upload(start.data(), start.size(), h.stream);
        h.exceptionOffsetsD.upload(entries.data(), entries.size(), h.stream);
        CUDA_CHECK(cudaStreamSynchronize(h.stream));
    }
    h.globalsD.alloc(std::max(h.G, 1));
    h.hGlobals.alloc(std::max(h.G, 1));
    h.selfD.alloc(2*SNB_MAX_SUBSETS);
    h.hSelf.alloc(2*SNB_MAX_SUBSETS);
    std::fill(h.hSelf.p, h.hSelf.p+2*SNB_MAX_SUBSETS, 0.0);
    if (!h.selfPartial.p) {
        h.selfPartial.alloc((size_t) SNB_PARAM_MAX_BLOCKS*2*SNB_MAX_SUBSETS);
        h.paramTicket.alloc(1);
        CUDA_CHECK(cudaMemsetAsync(h.paramTicket.p, 0, sizeof(unsigned), h.stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(h.stream));
}

void setupPme(snb_handle& h, PmeGrid& g, const int n[3], double alpha) {
    g.alpha = alpha;
    for (int k = 0; k < 3; k++) g.n[k] = n[k];
    size_t G = (size_t) n[0]*n[1]*n[2], GC = (size_t) n[0]*n[1]*(n[2]/2+1);
    const size_t rb = h.doublePrecision ? 8 : 4;
    g.real.alloc(G*h.S*rb);
    g.cplx.alloc(GC*h.S*2*rb);
    g.eterm.alloc(GC*rb);
    if (h.deterministic)
        g.fixed.alloc(G*h.S);
    // shared-memory bricks + TMA (pme_brick.cu): mixed precision, float atomics allowed (not the deterministic mode),
    // every dimension at least one brick long
    // Measured on B200 (gpurun_out r2q, profiles/README.md): the brick interpolation wins from a few hundred thousand atoms
    // on (STMV size: 146 us with its combine pass against 390 us) and loses below (ApoA1 size 134 vs 112 us, DHFR size 48 vs
    // 39 us: a CTA's chain of vote -> 64 KB bulk load -> 125 points is longer than the per-atom kernel's when there are too
    // few CTAs to overlap).  SNB_PME_BRICK_MIN_ATOMS overrides the threshold (A/B measurements).
    g.brick = false;
    const char* brickEnv = getenv("SNB_PME_BRICK_MIN_ATOMS");
    const long brickMinAtoms = brickEnv ? atol(brickEnv) : SNB_PME_BRICK_MIN_ATOMS;
    if (!h.doublePrecision && !h.deterministic && h.N >= brickMinAtoms && n[0] >= 20 && n[1] >= 20 && n[2] >= 24 && getenv("SNB_NO_PME_BRICKS") == nullptr) {
        g.nzp = (n[2]+4+3)&~3;
        g.pad.alloc((size_t) h.S*(n[0]+4)*(n[1]+4)*g.nzp);
        g.brick = makeBrickTensorMap(&g.brickMap, g.pad.p, h.N, h.S, n[0], n[1], g.nzp);
    }
    int dims[3] = {n[0], n[1], n[2]};
    CUFFT_CHECK(cufftPlanMany(&g.fwd, 3, dims, nullptr, 1, (int) G, nullptr, 1, (int) GC, h.doublePrecision ? CUFFT_D2Z : CUFFT_R2C, h.S));
    CUFFT_CHECK(cufftPlanMany(&g.inv, 3, dims, nullptr, 1, (int) GC, nullptr, 1, (int) G, h.doublePrecision ? CUFFT_Z2D : CUFFT_C2R, h.S));
    g.planned = true;
    CUFFT_CHECK(cufftSetStream(g.fwd, h.stream));
    CUFFT_CHECK(cufftSetStream(g.inv, h.stream));
    for (int k = 0; k < 3; k++) {
        std::vector<double> m = bsplineModuli(n[k]);
        g.moduli[k].upload(m.data(), m.size(), h.stream);
        CUDA_CHECK(cudaStreamSynchronize(h.stream));
    }
}

snb_handle* createHandle(const snb_desc& d) {
    validateDesc(d);
    snb_handle* hp = new snb_handle();
    snb_handle& h = *hp;
    try {
        h.device = d.device_index;
        h.doublePrecision = (d.flags&SNB_FLAG_DOUBLE_PRECISION) != 0;
        h.deterministic = h.doublePrecision || (d.flags&SNB_FLAG_DETERMINISTIC) != 0;
        h.directKernelFlags = d.flags&(SNB_FLAG_DIRECT_PACKED|SNB_FLAG_DIRECT_GENERAL);
        CUDA_CHECK(cudaSetDevice(h.device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, h.device));
        h.numSMs = prop.multiProcessorCount;
        // Stream priorities (SNB_STREAM_PRIORITY: 0 none [default], 1 reciprocal space + exceptions first, 2 the list /
        // direct-space stream first).  The kernels of the main stream fill the GPU for most of the evaluation (the direct-space
        // kernel takes every register of an SM); with equal priorities the short kernels of reciprocal space queue behind
        // their CTAs and the two chains run one after the other rather than side by side.
        int prioLow = 0, prioHigh = 0, mode = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
        if (const char* env = getenv("SNB_STREAM_PRIORITY")) mode = atoi(env);
        CUDA_CHECK(cudaStreamCreateWithPriority(&h.stream, cudaStreamNonBlocking, mode == 2 ? prioHigh : prioLow));
        CUDA_CHECK(cudaStreamCreateWithPriority(&h.pmeStream, cudaStreamNonBlocking, mode == 1 ? prioHigh : prioLow));
        CUDA_CHECK(cudaStreamCreateWithPriority(&h.bondedStream, cudaStreamNonBlocking, mode == 1 ? prioHigh : prioLow));
        CUDA_CHECK(cudaEventCreateWithFlags(&h.evInputs, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&h.evBondedDone, cudaEventDisableTiming));
        h.useGraph = getenv("SNB_NO_GRAPH") == nullptr;
        if (const char* env = getenv("SNB_FUSED_SORT")) h.fusedSort = atoi(env) != 0;
        if (const char* env = getenv("SNB_HOST_RESULTS")) h.hostResults = atoi(env) != 0;
        if (const char* env = getenv("SNB_BEGIN_KERNEL")) h.beginKernel = atoi(env) != 0;
        if (const char* env = getenv("SNB_COMPACT_SORT")) h.compactSort = atoi(env) != 0;
        h.stepD.alloc(1);
        h.stepH.alloc(1);
        memset(h.stepH.p, 0, sizeof(StepParams));
        {
            void* dev = nullptr;
            if (cudaHostGetDevicePointer(&dev, h.stepH.p, 0) == cudaSuccess) h.stepHDev = (const int*) dev;
            else cudaGetLastError();
        }
        CUDA_CHECK(cudaEventCreateWithFlags(&h.evSorted, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&h.evPmeDone, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreate(&h.evBegin));
        CUDA_CHECK(cudaEventCreate(&h.evEnd));
        h.N = d.num_particles;
        h.NB = (h.N+31)/32;
        h.NP = h.NB*32;
        h.S = d.num_subsets;
        h.L = h.S*(h.S+1)/2;
        h.G = d.num_global_parameters;
        h.X = d.num_exceptions;
        h.method = d.method;
        h.cutoff = d.cutoff;
        h.useSwitch = d.method != SNB_NO_CUTOFF && d.use_switching_function;
        h.switchDist = d.switching_distance;
        h.rfDielectric = d.reaction_field_dielectric;
        h.sliceParam.assign(h.L, {-1, -1});
        h.sliceHasDeriv.assign(h.L, {false, false});
        h.derivativeParams.assign(d.derivative_parameters, d.derivative_parameters+d.num_derivatives);
        std::set<int> requested(h.derivativeParams.begin(), h.derivativeParams.end());
        for (int k = 0; k < d.num_scaling_parameters; k++) {
            const int32_t* sp = d.scaling_parameters+5*k;
            if (sp[0] < 0 || sp[0] >= h.G || sp[1] < 0 || sp[1] >= h.S || sp[2] < 0 || sp[2] >= h.S)
                throw SnbException(SNB_ERR_INVALID, "scaling parameter out of range");
            int slice = snb_slice_index(sp[1], sp[2]);
            bool hasDeriv = requested.count(sp[0]) != 0;
            if (sp[3]) { h.sliceParam[slice][0] = sp[0]; h.sliceHasDeriv[slice][0] = hasDeriv; }
            if (sp[4]) { h.sliceParam[slice][1] = sp[0]; h.sliceHasDeriv[slice][1] = hasDeriv; }
        }
        loadParameters(h, d);
        h.is14 = find14s(d);
        h.num14 = (int) std::count(h.is14.begin(), h.is14.end(), 1);
        for (int k = 0; k < d.num_particle_offsets; k++) {
            h.particleOffsetIdx.push_back({d.particle_offset_index[2*k], d.particle_offset_index[2*k+1]});
            h.particleOffsetScale.push_back({d.particle_offset_scale[3*k], d.particle_offset_scale[3*k+1], d.particle_offset_scale[3*k+2]});
            h.offsetParams.insert(d.particle_offset_index[2*k]);
        }
        for (int k = 0; k < d.num_exception_offsets; k++) {
            h.exceptionOffsetIdx.push_back({d.exception_offset_index[2*k], d.exception_offset_index[2*k+1]});
            h.exceptionOffsetScale.push_back({d.exception_offset_scale[3*k], d.exception_offset_scale[3*k+1], d.exception_offset_scale[3*k+2]});
            h.offsetParams.insert(d.exception_offset_index[2*k]);
        }
        if (h.method == SNB_PME || h.method == SNB_LJPME)
            calcPMEParameters(d, false, h.alpha, h.grid);
        if (h.method == SNB_EWALD) {
            calcEwaldParameters(d, h.alpha, h.kmax);
            // the half space of reciprocal vectors the reference visits (ReferenceSlicedLJCoulombIxn.cpp:279-296)
            std::vector<int3> kv;
            int lowry = 0, lowrz = 1;
            for (int rx = 0; rx < h.kmax[0]; rx++) {
                for (int ry = lowry; ry < h.kmax[1]; ry++) {
                    for (int rz = lowrz; rz < h.kmax[2]; rz++)
                        kv.push_back(make_int3(rx, ry, rz));
                    lowrz = 1-h.kmax[2];
                }
                lowry = 1-h.kmax[1];
            }
            h.numK = (int) kv.size();
            h.ewaldK.alloc(std::max<size_t>(kv.size(), 1));
            if (!kv.empty())
                CUDA_CHECK(cudaMemcpy(h.ewaldK.p, kv.data(), sizeof(int3)*kv.size(), cudaMemcpyHostToDevice));
            h.ewaldS.alloc(std::max<size_t>(kv.size()*(size_t) h.S, 1));
        }
        if (h.method == SNB_LJPME) {
            calcPMEParameters(d, true, h.dalpha, h.dgrid);
            h.useSwitch = false;
        }
        h.exceptionsPeriodic = (h.method == SNB_NO_CUTOFF || h.method == SNB_CUTOFF_NON_PERIODIC) ? false : d.exceptions_use_pbc != 0;
        h.useDispersion = d.use_dispersion_correction != 0;
        h.dispersionCoef = h.useDispersion ? calcDispersionCorrections(d) : std::vector<double>(h.L, 0.0);

        // exclusion CSR over original indices (every exception excludes its pair, :104-112)
        std::vector<int> start(h.N+1, 0), list(2*(size_t) h.X);
        for (int e = 0; e < h.X; e++) {
            int a = h.exceptionAtoms[e].x, b = h.exceptionAtoms[e].y;
            if (a < 0 || a >= h.N || b < 0 || b >= h.N || a == b)
                throw SnbException(SNB_ERR_INVALID, "illegal particle index for an exception");
            start[a+1]++; start[b+1]++;
        }
        for (int i = 0; i < h.N; i++) start[i+1] += start[i];
        std::vector<int> fill(start.begin(), start.end()-1);
        for (int e = 0; e < h.X; e++) {
            int a = h.exceptionAtoms[e].x, b = h.exceptionAtoms[e].y;
            list[fill[a]++] = b;
            list[fill[b]++] = a;
        }
        h.exclStart.upload(start.data(), start.size(), h.stream);
        h.exclList.upload(list.data(), list.size(), h.stream);
        h.exceptionAtomsD.upload(h.exceptionAtoms.data(), h.X, h.stream);
        h.is14D.upload(h.is14.data(), h.X, h.stream);
        CUDA_CHECK(cudaStreamSynchronize(h.stream));

        h.posd.alloc(3*(size_t) h.N);
        h.paramsD.alloc(3*(size_t) h.N);
        h.params4.alloc(h.N);
        h.subsetD.alloc(h.N);
        h.exception14.alloc(3*(size_t) h.X);
        uploadBaseParameters(h);
        h.atomCell.alloc(h.N); h.atomSlot.alloc(h.N); h.cellScratch.alloc(h.N); h.slotCell.alloc(h.N);
        h.sortedAtom.alloc(h.NP); h.sortedPos.alloc(h.N);
        h.posqLocal.alloc(h.NP); h.sparams.alloc(h.NP);
        if (h.doublePrecision) { h.posqLocalD.alloc(h.NP); h.sparamsD.alloc(h.NP); }
        h.blockCenter.alloc(h.NB); h.blockExt.alloc(h.NB);
        h.octetCenter.alloc(4*(size_t) h.NB); h.octetExt.alloc(4*(size_t) h.NB);
        h.extent.alloc(6);
        // everything an evaluation must find zeroed lives in ONE allocation, cleared by one memset: the two
        // fixed-point force accumulators, the slice-energy table (32 replicas, summed on the host, + flags), the counters
        {
            const size_t f = 3*(size_t) h.NP*sizeof(long long), e = SNB_ENERGY_DOUBLES*sizeof(double);
            h.zeroBytes = 2*f+e+8*sizeof(int);
            h.zeroBlock.alloc(h.zeroBytes);
            h.forceSorted.view(h.zeroBlock.p, 3*(size_t) h.NP);
            h.forceOrig.view(h.zeroBlock.p+f, 3*(size_t) h.NP);
            h.energyBuf.view(h.zeroBlock.p+2*f, SNB_ENERGY_DOUBLES);
            h.counters.view(h.zeroBlock.p+2*f+e, 8);
        }
        h.forcesOutD.alloc(3*(size_t) h.N);
        h.hPos.alloc(3*(size_t) h.N); h.hForces.alloc(3*(size_t) h.N);
        h.hEnergy.alloc(SNB_ENERGY_DOUBLES);
        h.hCounters.alloc(8);
        // device-side addresses of the two pinned mirrors (unified addressing maps cudaMallocHost memory): k_finish stores
        // its results there; without a mapping the evaluation keeps its two device->host copies
        void *eDev = nullptr, *cDev = nullptr;
        if (cudaHostGetDevicePointer(&eDev, h.hEnergy.p, 0) == cudaSuccess && cudaHostGetDevicePointer(&cDev, h.hCounters.p, 0) == cudaSuccess) {
            h.hEnergyDev = (double*) eDev;
            h.hCountersDev = (int*) cDev;
        }
        else
            cudaGetLastError();
        h.sliceEnergies.assign(2*h.L, 0.0);
        if (h.method == SNB_PME || h.method == SNB_LJPME)
            setupPme(h, h.pme, h.grid, h.alpha);
        if (h.method == SNB_LJPME)
            setupPme(h, h.dpme, h.dgrid, h.dalpha);
    }
    catch (...) {
        delete hp;
        throw;
    }
    return hp;
}

void updateHandle(snb_handle& h, const snb_desc& d) {
    // ReferenceOpenMMLabKernels.cpp:263-312
    if (d.num_particles != h.N)
        throw SnbException(SNB_ERR_PARTICLE_COUNT, "updateParametersInContext: The number of particles has changed");
    validateDesc(d);
    std::vector<int> is14 = find14s(d);
    if ((int) std::count(is14.begin(), is14.end(), 1) != h.num14 || d.num_exceptions != h.X)
        throw SnbException(SNB_ERR_EXCEPTION_SET, "updateParametersInContext: The number of non-excluded exceptions has changed");
    // the exclusion structure was built from the exception pairs at creation: they must not move
    // (as the reference's CUDA platform insists, CudaOpenMMLabKernels.cpp:1175-1176)
    for (int e = 0; e < h.X; e++)
        if (d.exception_particles[2*e] != h.exceptionAtoms[e].x || d.exception_particles[2*e+1] != h.exceptionAtoms[e].y)
            throw SnbException(SNB_ERR_EXCEPTION_SET, "updateParametersInContext: The set of non-excluded exceptions has changed");
    CUDA_CHECK(cudaSetDevice(h.device));
    h.is14 = is14;
    h.is14D.upload(h.is14.data(), h.X, h.stream);
    loadParameters(h, d);
    CUDA_CHECK(cudaStreamSynchronize(h.stream));
    uploadBaseParameters(h);
    if (d.use_dispersion_correction && (d.method == SNB_CUTOFF_PERIODIC || d.method == SNB_EWALD || d.method == SNB_PME))
        h.dispersionCoef = calcDispersionCorrections(d);
}

// computeParameters (:332-385): the slice lambdas on the host (L values); particle and exception parameters with their
// offsets and the self energies on the device (params.cu), only when something they depend on changed, with no
// synchronisation: the kernel is ordered before the evaluation on the handle's stream, the self energies come back with
// an asynchronous copy that the evaluation's final synchronisation covers.
void computeParameters(snb_handle& h, const double* globals) {
    for (int s = 0; s < h.L; s++) {
        h.lambdaC[s] = h.sliceParam[s][0] < 0 ? 1.0 : globals[h.sliceParam[s][0]];
        h.lambdaV[s] = h.sliceParam[s][1] < 0 ? 1.0 : globals[h.sliceParam[s][1]];
    }
    std::vector<double> offsetValues;
    for (int p : h.offsetParams)
        offsetValues.push_back(globals[p]);
    if (!h.paramsDirty && offsetValues == h.lastOffsetValues)
        return;
    h.lastOffsetValues = offsetValues;
    h.paramsDirty = false;
    Range range("snb computeParameters");
    for (int g = 0; g < h.G; g++)
        h.hGlobals.p[g] = globals[g];
    if (h.G > 0)
        CUDA_CHECK(cudaMemcpyAsync(h.globalsD.p, h.hGlobals.p, sizeof(double)*h.G, cudaMemcpyHostToDevice, h.stream));
    ParamArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.numAtoms = h.N; pa.numExceptions = h.X;
    const bool ewaldFamily = h.method == SNB_EWALD || h.method == SNB_PME || h.method == SNB_LJPME;
    pa.selfScaleC = ewaldFamily ? SNB_ONE_4PI_EPS0*h.alpha/sqrt(M_PI) : 0.0;
    pa.selfScaleV = h.method == SNB_LJPME ? pow(h.dalpha, 6.0)*64.0/12.0 : 0.0;
    pa.globals = h.globalsD.p;
    pa.baseParticle = h.baseParticleD.p;
    pa.particleOffsetStart = h.particleOffsetIdx.empty() ? nullptr : h.particleOffsetStart.p;
    pa.particleOffsets = h.particleOffsetsD.p;
    pa.baseException = h.baseExceptionD.p;
    pa.exceptionOffsetStart = h.exceptionOffsetIdx.empty() ? nullptr : h.exceptionOffsetStart.p;
    pa.exceptionOffsets = h.exceptionOffsetsD.p;
    pa.subset = h.subsetD.p;
    pa.paramsD = h.paramsD.p; pa.params4 = h.params4.p; pa.exception14 = h.exception14.p;
    pa.selfPartial = h.selfPartial.p; pa.selfOut = h.selfD.p; pa.ticket = h.paramTicket.p;
    launchComputeParameters(pa, h.stream);
    CUDA_CHECK(cudaMemcpyAsync(h.hSelf.p, h.selfD.p, sizeof(double)*2*SNB_MAX_SUBSETS, cudaMemcpyDeviceToHost, h.stream));
    h.paramLaunches++;
}

void makeBox(const double* box, BoxD& bd, BoxF& bf) {
    bd.ax = box[0]; bd.bx = box[3]; bd.by = box[4]; bd.cx = box[6]; bd.cy = box[7]; bd.cz = box[8];
    double det = bd.ax*bd.by*bd.cz;
    memset(bd.recip, 0, sizeof(bd.recip));
    bd.recip[0][0] = bd.by*bd.cz/det;
    bd.recip[1][0] = -bd.bx*bd.cz/det;
    bd.recip[1][1] = bd.ax*bd.cz/det;
    bd.recip[2][0] = (bd.bx*bd.cy-bd.by*bd.cx)/det;
    bd.recip[2][1] = -bd.ax*bd.cy/det;
    bd.recip[2][2] = bd.ax*bd.by/det;
    bf.ax = (float) bd.ax; bf.bx = (float) bd.bx; bf.by = (float) bd.by;
    bf.cx = (float) bd.cx; bf.cy = (float) bd.cy; bf.cz = (float) bd.cz;
    bf.invAx = (float) (1.0/bd.ax); bf.invBy = (float) (1.0/bd.by); bf.invCz = (float) (1.0/bd.cz);
}

void setupCells(snb_handle& h, const double* box, bool periodic) {
    int dim[3];
    if (periodic) {
        double volume = box[0]*box[4]*box[8];
        double cs = cbrt(4.0*volume/std::max(h.N, 1));
        for (int k = 0; k < 3; k++)
            dim[k] = std::max(1, std::min(256, (int) floor(box[4*k]/cs+0.5)));
    }
    else {
        int m = std::max(1, std::min(256, (int) floor(cbrt(std::max(h.N, 1)/4.0)+0.5)));
        dim[0] = dim[1] = dim[2] = m;
    }
    if (dim[0] == h.cellDim[0] && dim[1] == h.cellDim[1] && dim[2] == h.cellDim[2])
        return;
    for (int k = 0; k < 3; k++) h.cellDim[k] = dim[k];
    h.numCells = dim[0]*dim[1]*dim[2];
    int bits = 1;
    while ((1<<bits) < std::max(dim[0], std::max(dim[1], dim[2]))) bits++;
    std::vector<std::pair<uint64_t, int>> keyed(h.numCells);
    for (int x = 0; x < dim[0]; x++)
        for (int y = 0; y < dim[1]; y++)
            for (int z = 0; z < dim[2]; z++) {
                int c = (x*dim[1]+y)*dim[2]+z;
                keyed[c] = {hilbertIndex(x, y, z, bits), c};
            }
    std::sort(keyed.begin(), keyed.end());
    std::vector<int> rank(h.numCells);
    for (int r = 0; r < h.numCells; r++)
        rank[keyed[r].second] = r;
    h.cellRank.upload(rank.data(), rank.size(), h.stream);
    CUDA_CHECK(cudaStreamSynchronize(h.stream));
    h.cellCount.alloc(h.numCells);
    CUDA_CHECK(cudaMemsetAsync(h.cellCount.p, 0, sizeof(int)*h.numCells, h.stream));    // k_scatterAtoms re-zeroes it after every use
    CUDA_CHECK(cudaStreamSynchronize(h.stream));
    h.cellStart.alloc(h.numCells+1);
    h.cellRange.alloc(h.numCells);
    h.tileSum.alloc(h.numCells/1024+2);
}

void ensureListCapacity(snb_handle& h, const double* box, bool periodic) {
    if (h.maxChunks > 0)
        return;
    // chunks of 32 j atoms per octet (the i side of the list works on octets of 8 atoms)
    size_t perOctet;
    if (h.method == SNB_NO_CUTOFF)
        perOctet = h.NB+2;
    else {
        double density = periodic ? h.N/(box[0]*box[4]*box[8]) : 100.0;
        double reach = h.cutoff+h.skinNow+0.3;
        double atoms = 0.5*density*(4.0/3.0*M_PI*reach*reach*reach)*1.3+72;
        perOctet = std::min<size_t>((size_t) (atoms/32)+2, (size_t) h.NB+2);
    }
    h.maxChunks = std::max<size_t>(64, perOctet*4*h.NB);
    h.jList.alloc(h.maxChunks*32);
    h.chunkOctet.alloc(h.maxChunks);
}

struct Timer {
    snb_handle& h;
    explicit Timer(snb_handle& h) : h(h) {
        if (h.timing && !h.evCreated) {
            for (auto& e : h.ev) CUDA_CHECK(cudaEventCreate(&e));
            h.evCreated = true;
        }
    }
    void mark(int k) { if (h.timing) CUDA_CHECK(cudaEventRecord(h.ev[k], h.stream)); }
};

// Rank 0's share of the tile list relative to 1.0 for every other rank.  Cost model: reciprocal space
// (rank 0 only) costs about rho = 0.7 of the whole direct-space phase of the same system on one GPU, so
// balancing  P + f0 D = (1 - f0) D / (W - 1)  gives f0 = (1 - (W-1) rho) / W of the octets to rank 0.
double rank0Share(const snb_handle& h) {
    if (h.world <= 1) return 1.0;
    if (h.rank0Share >= 0) return h.rank0Share;
    if (const char* env = getenv("SNB_RANK0_SHARE")) return std::max(0.0, atof(env));
    if (h.tunedShare >= 0) return h.tunedShare;         // measured (runOnce): same value on every rank
    if (h.method != SNB_PME && h.method != SNB_LJPME) return 1.0;
    const double rho = h.method == SNB_LJPME ? 0.8 : 0.4;
    const double f0 = std::max(0.0, (1.0-(h.world-1)*rho)/h.world);
    return f0*(h.world-1)/(1.0-f0);
}

struct FusedSortRefused { std::string msg; };      // the cooperative launch of k_sortFused was refused (nothing has run): fall back

struct StepInputs {
    const double* hostPos;      // host positions to copy from: the caller's own buffer when it is page-locked, else the pinned mirror
    const void* posqDev;        // device float4 positions, or null when the positions come from the host
    void* forcesFixedOut;
    int flags;
    bool hostPositions;
    int pbcMode;
};

// Host-side preparation of one evaluation: cells, capacities, StepParams mirror, reciprocal-space table.
int prepareStep(snb_handle& h, const double* box, bool periodic) {
    BoxD bd;
    BoxF bf;
    makeBox(box, bd, bf);
    h.skinNow = h.method == SNB_NO_CUTOFF ? 0.0 : h.listSkin;
    if (periodic && h.skinNow > 0) {
        const double room = 0.5*std::min(bd.ax, std::min(bd.by, bd.cz))-h.cutoff;
        h.skinNow = std::max(0.0, std::min(h.skinNow, 0.9*room));
    }
    setupCells(h, box, periodic);
    ensureListCapacity(h, box, periodic);
    int pbcMode = PBC_NONE;
    if (periodic) {
        const bool triclinic = bd.bx != 0 || bd.cx != 0 || bd.cy != 0;
        pbcMode = triclinic ? PBC_PAIR_TRICLINIC : PBC_SHIFT;   // PBC_SHIFT: per work item, one image per j atom or per-pair wrapping
    }
    StepParams& sp = *h.stepH.p;
    sp.exact.boxd = bd;
    sp.exact.cutoff2d = h.cutoff*h.cutoff;
    sp.exact.periodic = periodic;
    sp.box = bf;
    sp.shiftLimit[0] = (float) (0.5*bd.ax-h.cutoff); sp.shiftLimit[1] = (float) (0.5*bd.by-h.cutoff); sp.shiftLimit[2] = (float) (0.5*bd.cz-h.cutoff);
    for (int s = 0; s < SNB_MAX_SLICES; s++) {
        sp.lambdaC[s] = s < h.L ? h.lambdaC[s] : 0.0;
        sp.lambdaV[s] = s < h.L ? h.lambdaV[s] : 0.0;
    }
    // list reuse: the host's own reasons for a rebuild (the device adds "an atom has moved more than skin/2", sort.cu)
    sp.skinHalf2 = 0.25*h.skinNow*h.skinNow;
    sp.forceRebuild = 0;
    if (h.skinNow > 0) {
        const double share = h.comm ? rank0Share(h) : 0.0;
        if (!h.listKept || memcmp(h.listBox, box, sizeof(h.listBox)) != 0 || h.listMaxChunks != h.maxChunks || h.listShare != share || h.listBuiltSkin != h.skinNow)
            sp.forceRebuild = 1;
    }
    return pbcMode;
}

PmeArgs pmeArgs(snb_handle& h, int term, bool sorted) {
    PmeGrid& g = term == 0 ? h.pme : h.dpme;
    PmeArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.numAtoms = h.N; pa.numSubsets = h.S; pa.term = term;
    pa.nx = g.n[0]; pa.ny = g.n[1]; pa.nz = g.n[2];
    pa.alpha = g.alpha; pa.sp = h.stepD.p; pa.posd = h.posd.p; pa.doublePrecision = h.doublePrecision;
    pa.sortedAtom = sorted ? h.sortedAtom.p : nullptr;
    pa.params4 = h.params4.p; pa.paramsD = h.paramsD.p; pa.grid = g.real.p; pa.cgrid = g.cplx.p; pa.eterm = g.eterm.p;
    pa.gridFixed = h.deterministic ? g.fixed.p : nullptr;
    pa.gridPad = g.brick ? g.pad.p : nullptr; pa.nzp = g.nzp;
    for (int k = 0; k < 3; k++) pa.moduli[k] = g.moduli[k].p;
    pa.forceOrig = h.forceOrig.p; pa.paddedAtoms = h.NP; pa.energyBuf = h.energyBuf.p;
    return pa;
}

template <typename R> void fillConsts(DirectConsts<R>& c, const DirectParams& p) {
    c.cutoffHi = (R) (p.cutoff2d+p.bandTol); c.cutoffLo = (R) (p.cutoff2d-p.bandTol);
    c.cutoffHiWide = (R) (p.cutoff2d+p.bandTolWide); c.cutoffLoWide = (R) (p.cutoff2d-p.bandTolWide);
    c.alpha = (R) p.alphad; c.alpha2 = (R) (p.alphad*p.alphad); c.halfAlpha = (R) (0.5*p.alphad);
    c.negAlpha2Log2e = (R) (-p.alphad*p.alphad*1.4426950408889634);
    c.twoAlphaOverSqrtPi = (R) (2.0*p.alphad/1.772453850905516);
    c.krf = (R) p.krfd; c.crf = (R) p.crfd; c.twoKrf = (R) (2.0*p.krfd);
    // no switching function: the test `r > switchDist` must never fire
    c.switchDist = p.invSwitchWidthd != 0.0 ? (R) p.switchDistd : (R) INFINITY; c.invSwitchWidth = (R) p.invSwitchWidthd;
    c.dalpha2 = (R) (p.dalphad*p.dalphad); c.dispShiftExp = (R) p.dispShiftExpd; c.invCut6 = (R) p.invCut6d;
}

// Enqueue one whole evaluation on the handle's streams (no host synchronisation, no allocation: the
// sequence is capturable as a CUDA graph).
void enqueueStep(snb_handle& h, const StepInputs& in) {
    const int flags = in.flags;
    const bool includeForces = flags&SNB_INCLUDE_FORCES, includeDirect = flags&SNB_INCLUDE_DIRECT,
               includeRecip = flags&SNB_INCLUDE_RECIPROCAL;
    const bool periodic = h.method >= SNB_CUTOFF_PERIODIC;
    const bool pmeFamily = h.method == SNB_PME || h.method == SNB_LJPME;
    const bool ewaldFamily = pmeFamily || h.method == SNB_EWALD;
    const bool needEnergy = (flags&SNB_INCLUDE_ENERGY) || !h.derivativeParams.empty();
    cudaStream_t st = h.stream;
    const NcclApi* nccl = h.comm ? ncclApi(nullptr) : nullptr;
    Timer tm(h);

    // ---- inputs ----
    // multi-GPU with the position broadcast on: rank 0's positions are authoritative -- ONE host->device copy on rank 0,
    // the other ranks receive them over NVLink -- and every rank then builds the identical sorted order
    const bool root = h.rank == 0;
    const bool receivePositions = nccl && h.broadcastPositions && !root;
    SortArgs sa;
    memset(&sa, 0, sizeof(sa));
    if (includeDirect) {
        sa.numAtoms = h.N; sa.paddedAtoms = h.NP; sa.numBlocks = h.NB; sa.numCells = h.numCells; sa.periodic = periodic;
        for (int k = 0; k < 3; k++) sa.cellDim[k] = h.cellDim[k];
        sa.counters = h.counters.p;
        sa.sqrtK = (float) sqrt(SNB_ONE_4PI_EPS0);
        sa.sp = h.stepD.p;
        sa.posd = h.posd.p; sa.params4 = h.params4.p; sa.paramsD = h.paramsD.p; sa.cellRank = h.cellRank.p; sa.extent = h.extent.p;
        sa.posqLocalD = h.doublePrecision ? h.posqLocalD.p : nullptr; sa.sparamsD = h.doublePrecision ? h.sparamsD.p : nullptr;
        sa.cellRange = h.cellRange.p; sa.tileSum = h.tileSum.p; sa.cellCount = h.cellCount.p; sa.cellStart = h.cellStart.p; sa.atomCell = h.atomCell.p; sa.atomSlot = h.atomSlot.p; sa.cellScratch = h.cellScratch.p;
        sa.sortedAtom = h.sortedAtom.p; sa.sortedPos = h.sortedPos.p; sa.posqLocal = h.posqLocal.p; sa.sparams = h.sparams.p;
        sa.blockCenter = h.blockCenter.p; sa.blockExt = h.blockExt.p; sa.octetCenter = h.octetCenter.p; sa.octetExt = h.octetExt.p;
        sa.reuse = h.skinNow > 0; sa.posRef = h.posRef.p; sa.wrapImage = h.wrapImage.p;
        // (the cooperative kernel stays out of multi-GPU evaluations: untested beside NCCL's kernels)
        sa.fused = h.fusedSort && !h.comm; sa.numSMs = h.numSMs; sa.compact = h.compactSort; sa.slotCell = h.slotCell.p;
    }
    // single GPU: the copy of the step's parameters, the memset of the accumulators and the position conversion are ONE
    // kernel (k_begin, finish.cu) -- three graph nodes fewer at the head of every chain of the evaluation
    const bool beginKernel = !nccl && h.beginKernel && h.stepHDev != nullptr;
    h.beginKernelUsed = beginKernel;
    Range* range = new Range("snb inputs");
    if (in.hostPositions) {
        if (!receivePositions)
            CUDA_CHECK(cudaMemcpyAsync(h.posd.p, in.hostPos, sizeof(double)*3*(size_t) h.N, cudaMemcpyHostToDevice, st));
    }
    else if (!receivePositions && !beginKernel)
        launchConvertPositions(in.posqDev, h.ordCorrection, h.ordAtomIndex, h.ordDouble, h.posd.p, h.N, st);
    if (beginKernel) {
        BeginArgs bg;
        memset(&bg, 0, sizeof(bg));
        bg.stepHost = h.stepHDev; bg.stepDev = (int*) h.stepD.p; bg.stepWords = (int) (sizeof(StepParams)/4);
        bg.zero = (unsigned long long*) h.zeroBlock.p; bg.zeroWords = h.zeroBytes/8;
        bg.posq = in.hostPositions ? nullptr : in.posqDev; bg.correction = (const float4*) h.ordCorrection; bg.atomIndex = h.ordAtomIndex;
        bg.positionsDouble = h.ordDouble; bg.posd = h.posd.p; bg.numAtoms = h.N;
        launchBegin(bg, st);
    }
    if (nccl && h.broadcastPositions)
        NCCL_CHECK(nccl, nccl->Broadcast(h.posd.p, h.posd.p, 3*(size_t) h.N, ncclFloat64, 0, h.comm, st));
    if (!beginKernel) {
        CUDA_CHECK(cudaMemcpyAsync(h.stepD.p, h.stepH.p, sizeof(StepParams), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemsetAsync(h.zeroBlock.p, 0, h.zeroBytes, st));
    }
    // multi-GPU, rank 0: reciprocal space accumulates into buffers of its own, so that the other terms can be reduced
    // between the ranks WHILE it runs; its forces and energies are added after the reduction
    const bool splitRecip = nccl && root;
    if (splitRecip) {
        CUDA_CHECK(cudaMemsetAsync(h.forcePme.p, 0, sizeof(long long)*3*(size_t) h.NP, st));
        CUDA_CHECK(cudaMemsetAsync(h.energyPme.p, 0, sizeof(double)*2*SNB_MAX_SLICES, st));
    }
    long long* recipForce = splitRecip ? h.forcePme.p : h.forceOrig.p;
    double* recipEnergy = splitRecip ? h.energyPme.p : h.energyBuf.p;
    delete range;
    tm.mark(0);

    // ---- exceptions / exclusion corrections: need the inputs only, so they run beside the sort ----
    bool bondedForked = false;
    auto bonded = [&](cudaStream_t bs) {
        BondedArgs ba;
        memset(&ba, 0, sizeof(ba));
        int64_t xlo = 0, xhi = h.X;
        partitionRange(h.X, h.world, h.rank, 1.0, &xlo, &xhi);
        ba.firstException = (int) xlo;
        ba.numExceptions = (int) (xhi-xlo); ba.periodicExceptions = h.exceptionsPeriodic; ba.ewald = ewaldFamily; ba.ljpme = h.method == SNB_LJPME;
        ba.includeForces = 1;
        ba.alpha = h.alpha; ba.dalpha = h.dalpha; ba.sp = h.stepD.p;
        ba.posd = h.posd.p; ba.paramsD = h.paramsD.p; ba.subset = h.subsetD.p; ba.exceptionAtoms = h.exceptionAtomsD.p;
        ba.exception14 = h.exception14.p; ba.exceptionIs14 = h.is14D.p;
        ba.forceOrig = h.forceOrig.p; ba.paddedAtoms = h.NP; ba.energyBuf = h.energyBuf.p;
        launchBonded(ba, bs);
    };
    const bool inputsRecorded = !h.timing && ((includeDirect && h.X > 0) || h.pmeEarly || (includeRecip && pmeFamily && h.rank == 0));
    if (inputsRecorded)
        CUDA_CHECK(cudaEventRecord(h.evInputs, st));
    auto forkBonded = [&](cudaEvent_t after) {
        CUDA_CHECK(cudaStreamWaitEvent(h.bondedStream, after, 0));
        bonded(h.bondedStream);
        CUDA_CHECK(cudaEventRecord(h.evBondedDone, h.bondedStream));
        bondedForked = true;
    };
    if (includeDirect && h.X > 0 && !h.timing)
        forkBonded(h.evInputs);

    DirectArgs da;
    memset(&da, 0, sizeof(da));
    h.fusedSortUsed = false;
    h.pmeFoldedLaunches = 0;
    if (includeDirect) {
        Range sortRange("snb sort");
        bool used = false;
        const cudaError_t err = launchSort(sa, st, &used);
        h.fusedSortUsed = used;
        if (used && err != cudaSuccess)
            throw FusedSortRefused{std::string("cudaLaunchCooperativeKernel(k_sortFused): ")+cudaGetErrorString(err)};
        CUDA_CHECK(err);
    }
    tm.mark(1);
    if (!h.timing) CUDA_CHECK(cudaEventRecord(h.evSorted, st));

    // ---- reciprocal space (rank 0): its own stream, beside the tile list / direct-space kernels ----
    bool gridsZeroed = false;       // the spreading grids were cleared ahead of the sort's end (below)
    auto reciprocal = [&](cudaStream_t ps) {
        Range recipRange("snb reciprocal space");
        if (h.method == SNB_EWALD) {
            EwaldArgs ea;
            memset(&ea, 0, sizeof(ea));
            ea.numAtoms = h.N; ea.numSubsets = h.S; ea.numK = h.numK; ea.paddedAtoms = h.NP; ea.alpha = h.alpha;
            ea.sp = h.stepD.p; ea.posd = h.posd.p; ea.paramsD = h.paramsD.p; ea.subset = h.subsetD.p;
            ea.kvec = h.ewaldK.p; ea.structure = h.ewaldS.p; ea.forceOrig = recipForce;
            ea.energyBuf = needEnergy ? recipEnergy : nullptr;
            launchEwaldReciprocal(ea, includeForces, ps);
            tm.mark(5); tm.mark(6); tm.mark(7); tm.mark(8);
            return;
        }
        for (int term = 0; term < (h.method == SNB_LJPME ? 2 : 1); term++) {
            PmeGrid& g = term == 0 ? h.pme : h.dpme;
            PmeArgs pa = pmeArgs(h, term, includeDirect);
            if (h.pmeEarly && ps != st)
                pa.sortedAtom = h.sortedAtomPme.p;
            pa.forceOrig = recipForce; pa.energyBuf = recipEnergy;
            CUFFT_CHECK(cufftSetStream(g.fwd, ps));
            CUFFT_CHECK(cufftSetStream(g.inv, ps));
            // atoms in sorted order + mixed precision: shared-memory bricks moved by TMA; else the per-atom kernels
            const bool bricks = g.brick && pa.sortedAtom != nullptr;
            static const bool noBrickInterp = getenv("SNB_NO_BRICK_INTERP") != nullptr;
            launchPmeSpread(pa, ps, gridsZeroed);
            if (term == 0) tm.mark(5);
            if (h.doublePrecision) CUFFT_CHECK(cufftExecD2Z(g.fwd, (double*) g.real.p, (cufftDoubleComplex*) g.cplx.p));
            else CUFFT_CHECK(cufftExecR2C(g.fwd, (float*) g.real.p, (cufftComplex*) g.cplx.p));
            if (term == 0) tm.mark(6);
            launchPmeConvolution(pa, ps);
            if (term == 0) tm.mark(7);
            if (h.doublePrecision) CUFFT_CHECK(cufftExecZ2D(g.inv, (cufftDoubleComplex*) g.cplx.p, (double*) g.real.p));
            else CUFFT_CHECK(cufftExecC2R(g.inv, (cufftComplex*) g.cplx.p, (float*) g.real.p));
            if (bricks && !noBrickInterp) CUDA_CHECK(launchPmeInterpolateBrick(pa, g.brickMap, ps));
            else {
                launchPmeInterpolate(pa, ps);
                if (pmeInterpolateFolds(pa)) h.pmeFoldedLaunches++;
            }
            if (term == 0) tm.mark(8);
        }
    };
    const bool doRecip = includeRecip && ewaldFamily && h.rank == 0;
    if (doRecip && !h.timing) {
        static const bool earlyZeroOff = getenv("SNB_EARLY_GRID_ZERO") != nullptr && atoi(getenv("SNB_EARLY_GRID_ZERO")) == 0;
        // (small grids only: at STMV size the 60 MB memset ahead of the sort's end lets the spreading start together with the
        // list build, which then loses the SMs it needs first -- 3.01 vs 2.73 ms per evaluation, gpurun r3q)
        size_t gridBytes = 0;
        for (int term = 0; term < (h.method == SNB_LJPME ? 2 : 1); term++) {
            const PmeGrid& g = term == 0 ? h.pme : h.dpme;
            gridBytes += (size_t) h.S*g.n[0]*g.n[1]*g.n[2]*(h.doublePrecision ? 8 : 4);
        }
        if (pmeFamily && !h.pmeEarly && inputsRecorded && !earlyZeroOff && gridBytes <= (size_t) 16<<20) {
            // the grids are free since the previous evaluation: their memsets run beside the sort instead of behind it
            CUDA_CHECK(cudaStreamWaitEvent(h.pmeStream, h.evInputs, 0));
            for (int term = 0; term < (h.method == SNB_LJPME ? 2 : 1); term++) {
                PmeArgs pa = pmeArgs(h, term, includeDirect);
                launchPmeZero(pa, h.pmeStream);
            }
            gridsZeroed = true;
        }
        CUDA_CHECK(cudaStreamWaitEvent(h.pmeStream, h.pmeEarly ? h.evInputs : h.evSorted, 0));
        if (nccl && !h.balanceFrozen) CUDA_CHECK(cudaEventRecord(h.evRecBegin, h.pmeStream));
        reciprocal(h.pmeStream);
        if (nccl && !h.balanceFrozen) CUDA_CHECK(cudaEventRecord(h.evRecEnd, h.pmeStream));
        CUDA_CHECK(cudaEventRecord(h.evPmeDone, h.pmeStream));
    }
    if (nccl && includeDirect && !h.balanceFrozen) CUDA_CHECK(cudaEventRecord(h.evDirBegin, st));

    if (includeDirect) {
        Range directRange("snb list + direct + exceptions");
        ListArgs la;
        memset(&la, 0, sizeof(la));
        la.numAtoms = h.N; la.numBlocks = h.NB; la.pbcMode = in.pbcMode; la.periodic = periodic;
        la.useCutoff = h.method != SNB_NO_CUTOFF;
        la.cutoff = (float) ((h.cutoff+h.skinNow)*(1.0+1e-5)+1e-5);
        la.reuse = h.skinNow > 0; la.skin = (float) h.skinNow;
        la.sp = h.stepD.p;
        la.sortedAtom = h.sortedAtom.p; la.sortedPos = h.sortedPos.p; la.posqLocal = h.posqLocal.p;
        la.blockCenter = h.blockCenter.p; la.blockExt = h.blockExt.p; la.octetCenter = h.octetCenter.p; la.octetExt = h.octetExt.p;
        la.exclStart = h.exclStart.p; la.exclList = h.exclList.p;
        la.jList = h.jList.p; la.chunkOctet = h.chunkOctet.p; la.counters = h.counters.p;
        la.maxChunks = (int) h.maxChunks;
        la.cellRange = h.cellRange.p; la.extent = h.extent.p;
        for (int k = 0; k < 3; k++) la.cellDim[k] = h.cellDim[k];
        int64_t lo = 0, hi = h.NB;
        partitionRange(h.NB, h.world, h.rank, rank0Share(h), &lo, &hi);
        la.blockLo = (int) lo; la.blockHi = (int) hi;
        la.bitmapWords = (h.NB+31)/32;
        // Atom-level refinement of the list (41 % -> 55 % of the pair slots in range): the direct-space kernel gets 16-20 %
        // shorter and the list build 40-50 % longer.  Measured on B200: a wash for the whole evaluation at DHFR size (162.7 vs
        // 163.6 us), a loss at STMV size (3.05 vs 2.85 ms), so it is on below SNB_PACKED_MIN_ATOMS atoms only.
        // SNB_LIST_REFINE=0/1 overrides (A/B measurements).
        static const char* refineEnv = getenv("SNB_LIST_REFINE");
        // (a kept list pays for its build once in ~10 evaluations and for every entry in all of them: always refined)
        la.refine = refineEnv ? atoi(refineEnv) != 0 : (h.N < SNB_PACKED_MIN_ATOMS || h.skinNow > 0);
        CUDA_CHECK(launchBuildList(la, st, h.numSMs));
        tm.mark(2);

        DirectParams& p = da.p;
        p.numAtoms = h.N; p.numBlocks = h.NB; p.pbcMode = in.pbcMode; p.periodic = periodic;
        p.cutoff2d = h.cutoff*h.cutoff;
        p.bandTolWide = (float) (4e-5*h.cutoff*h.cutoff);
        p.bandTol = (float) fmin((double) p.bandTolWide, 2.0e-6*h.cutoff+1.0e-6*h.cutoff*h.cutoff);    // ~4x the error bound
        p.alphad = h.alpha;
        p.krfd = pow(h.cutoff, -3.0)*(h.rfDielectric-1.0)/(2.0*h.rfDielectric+1.0);
        p.crfd = (1.0/h.cutoff)*(3.0*h.rfDielectric)/(2.0*h.rfDielectric+1.0);
        p.switchDistd = h.switchDist;
        p.invSwitchWidthd = h.useSwitch ? 1.0/(h.cutoff-h.switchDist) : 0.0;
        double dc2 = h.dalpha*h.cutoff*h.dalpha*h.cutoff;
        p.dalphad = h.dalpha;
        p.dispShiftExpd = 1.0-exp(-dc2)*(1.0+dc2+0.5*dc2*dc2);
        p.invCut6d = pow(h.cutoff, -6.0);
        fillConsts(da.cf, p);
        fillConsts(da.cd, p);
        da.sp = h.stepD.p;
        da.kind = h.method == SNB_NO_CUTOFF ? KIND_PLAIN : h.method <= SNB_CUTOFF_PERIODIC ? KIND_RF : h.method == SNB_LJPME ? KIND_LJPME : KIND_EWALD;
        da.useSwitch = h.useSwitch; da.numSubsets = h.S; da.needEnergy = needEnergy; da.doublePrecision = h.doublePrecision;
        // small systems keep the general kernel: an evaluation of a few ten thousand atoms is latency bound, and sixteen
        // resident warps per SM finish a short chunk list sooner than the packed kernel's twelve (DHFR size: 43 vs 52 us)
        da.usePacked = (h.directKernelFlags&SNB_FLAG_DIRECT_PACKED) ? 1 : (h.directKernelFlags&SNB_FLAG_DIRECT_GENERAL) ? 0 : h.N >= SNB_PACKED_MIN_ATOMS;
        da.posqLocalD = h.posqLocalD.p; da.sparamsD = h.sparamsD.p;
        da.sortedAtom = h.sortedAtom.p; da.posqLocal = h.posqLocal.p; da.sparams = h.sparams.p;
        da.blockCenter = h.blockCenter.p; da.octetCenter = h.octetCenter.p; da.posd = h.posd.p; da.jList = h.jList.p; da.chunkOctet = h.chunkOctet.p; da.maxChunks = (int) h.maxChunks;
        da.counters = h.counters.p; da.listState = h.skinNow > 0 ? h.listState.p : nullptr; da.forceSorted = h.forceSorted.p; da.paddedAtoms = h.NP; da.energyBuf = h.energyBuf.p;
        h.packedLaunched = launchDirect(da, st, h.numSMs);
        if (nccl && !h.balanceFrozen) CUDA_CHECK(cudaEventRecord(h.evDirEnd, st));
        tm.mark(3);
        if (h.X > 0 && !bondedForked)
            bonded(st);
        tm.mark(4);
    }
    else {
        tm.mark(2); tm.mark(3); tm.mark(4);
    }
    if (doRecip && h.timing)
        reciprocal(st);
    else if (h.timing) {
        tm.mark(5); tm.mark(6); tm.mark(7); tm.mark(8);
    }
    if (bondedForked)
        CUDA_CHECK(cudaStreamWaitEvent(st, h.evBondedDone, 0));
    Range outRange("snb reduce + outputs");

    // ---- outputs ----
    bool resultsStored = false;     // k_finish wrote energies + counters into the pinned mirrors itself
    FinishArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.numAtoms = h.N; fa.paddedAtoms = h.NP;
    fa.sortedPos = includeDirect ? h.sortedPos.p : nullptr;
    fa.forceSorted = h.forceSorted.p; fa.forceOrig = h.forceOrig.p;
    fa.counters = h.counters.p;
    fa.outIndex = h.ordAtomIndex; fa.outStride = h.ordStride ? h.ordStride : h.N;
    if (nccl) {
        // Direct space + exceptions of every rank -> rank 0: ncclReduce(sum, int64) of the [3][N] fixed-point forces, grouped
        // with a small all-reduce of the slice-energy table, the list-overflow flag (every rank must take the same retry
        // decision) and the times every rank measured in its previous evaluation (every rank derives the same new partition),
        // all in 2^32 fixed point.  Integer sums are exact: the forces are bit-identical to the single-GPU ones whatever the
        // partition.  Rank 0's reciprocal space is NOT waited for first: it runs on its own stream while the collective is in
        // flight, and its forces / energies are added afterwards (reciprocal space stays on rank 0, as in the reference:
        // platforms/cuda/src/CudaOpenMMLabKernels.cpp:379,405,466).
        if (!h.balanceFrozen) CUDA_CHECK(cudaEventRecord(h.evReady, st));
        if (h.loadValid && !h.balanceFrozen)
            CUDA_CHECK(cudaMemcpyAsync(h.energyBuf.p+32*2*SNB_MAX_SLICES+1, h.loadH.p+1, sizeof(double)*(SNB_ENERGY_FLAGS-1), cudaMemcpyHostToDevice, st));
        fa.reduceOut = h.reduceSend.p;
        fa.energyIn = h.energyBuf.p; fa.numEnergy = SNB_ENERGY_DOUBLES; fa.overflowSlot = 32*2*SNB_MAX_SLICES;
        launchFinish(fa, st);
        const size_t nf = 3*(size_t) h.N;
        NCCL_CHECK(nccl, nccl->GroupStart());
        NCCL_CHECK(nccl, nccl->Reduce(h.reduceSend.p, h.reduceRecv.p, nf, ncclInt64, ncclSum, 0, h.comm, st));
        NCCL_CHECK(nccl, nccl->AllReduce(h.reduceSend.p+nf, h.reduceRecv.p+nf, SNB_ENERGY_DOUBLES, ncclInt64, ncclSum, h.comm, st));
        NCCL_CHECK(nccl, nccl->GroupEnd());
        if (doRecip && !h.timing)
            CUDA_CHECK(cudaStreamWaitEvent(st, h.evPmeDone, 0));
        // (reciprocal space is done with the copy of the previous order; this evaluation's order is final: next copy)
        if (root && includeDirect && h.sortedAtomPme.p)
            CUDA_CHECK(cudaMemcpyAsync(h.sortedAtomPme.p, h.sortedAtom.p, sizeof(int)*(size_t) h.NP, cudaMemcpyDeviceToDevice, st));
        if (includeForces && root)
            launchEmitReduced(h.reduceRecv.p, h.N, h.NP, splitRecip ? h.forcePme.p : nullptr, in.hostPositions ? h.forcesOutD.p : nullptr,
                              (long long*) in.forcesFixedOut, h.ordAtomIndex, h.ordStride ? h.ordStride : h.N, h.reduceRecv.p+nf+32*2*SNB_MAX_SLICES, st);
    }
    else {
        if (doRecip && !h.timing)
            CUDA_CHECK(cudaStreamWaitEvent(st, h.evPmeDone, 0));
        if (includeForces) {
            fa.forcesOut = in.hostPositions ? h.forcesOutD.p : nullptr;
            fa.forcesFixedOut = (long long*) in.forcesFixedOut;
            if (h.hostResults && h.hEnergyDev && h.hCountersDev) {
                fa.hostEnergy = h.hEnergyDev; fa.energySrc = h.energyBuf.p; fa.numEnergyHost = SNB_ENERGY_DOUBLES; fa.hostCounters = h.hCountersDev;
                resultsStored = true;
            }
            launchFinish(fa, st);
        }
    }
    if (includeForces && in.hostPositions && (!nccl || root))
        CUDA_CHECK(cudaMemcpyAsync(h.hForces.p, h.forcesOutD.p, sizeof(double)*3*(size_t) h.N, cudaMemcpyDeviceToHost, st));
    tm.mark(9);
    if (nccl) {
        CUDA_CHECK(cudaMemcpyAsync(h.hEnergyFixed.p, h.reduceRecv.p+3*(size_t) h.N, sizeof(long long)*SNB_ENERGY_DOUBLES, cudaMemcpyDeviceToHost, st));
        if (splitRecip)
            CUDA_CHECK(cudaMemcpyAsync(h.hEnergyPme.p, h.energyPme.p, sizeof(double)*2*SNB_MAX_SLICES, cudaMemcpyDeviceToHost, st));
    }
    else if (!resultsStored)
        CUDA_CHECK(cudaMemcpyAsync(h.hEnergy.p, h.energyBuf.p, sizeof(double)*SNB_ENERGY_DOUBLES, cudaMemcpyDeviceToHost, st));
    if (!resultsStored)
        CUDA_CHECK(cudaMemcpyAsync(h.hCounters.p, h.counters.p, sizeof(int)*8, cudaMemcpyDeviceToHost, st));
    h.lastDirect = da;
}

void dropGraph(snb_handle& h) {
    if (h.graphExec) cudaGraphExecDestroy(h.graphExec);
    h.graphExec = nullptr;
    h.graphKeyValid = false;
}

// One evaluation.  Returns false when a device-side capacity check asks for a repeat (buffers were grown).
bool runOnce(snb_handle& h, const double* box, int flags, const double* hostPos, const void* posqDev, void* forcesFixedOut) {
    const bool includeDirect = flags&SNB_INCLUDE_DIRECT, includeRecip = flags&SNB_INCLUDE_RECIPROCAL, includeForces = flags&SNB_INCLUDE_FORCES;
    const bool periodic = h.method >= SNB_CUTOFF_PERIODIC;
    const bool pmeFamily = h.method == SNB_PME || h.method == SNB_LJPME;
    cudaStream_t st = h.stream;
    StepInputs in;
    in.hostPos = hostPos; in.posqDev = posqDev; in.forcesFixedOut = forcesFixedOut; in.flags = flags; in.hostPositions = posqDev == nullptr;
    in.pbcMode = prepareStep(h, box, periodic);

    // reciprocal-space kernel table: rebuilt only when the box changes (outside the graph, same stream)
    if (includeRecip && pmeFamily && h.rank == 0)
        for (int term = 0; term < (h.method == SNB_LJPME ? 2 : 1); term++) {
            PmeGrid& g = term == 0 ? h.pme : h.dpme;
            const BoxD& bd = h.stepH.p->exact.boxd;
            double key[6] = {bd.ax, bd.bx, bd.by, bd.cx, bd.cy, bd.cz};
            if (!g.etermValid || memcmp(key, g.etermBox, sizeof(key)) != 0) {
                // the kernel reads the box from StepParams: refresh the device copy first
                CUDA_CHECK(cudaMemcpyAsync(h.stepD.p, h.stepH.p, sizeof(StepParams), cudaMemcpyHostToDevice, st));
                launchPmeEterm(pmeArgs(h, term, false), st);
                memcpy(g.etermBox, key, sizeof(key));
                g.etermValid = true;
            }
        }

    h.pmeEarly = h.comm && h.rank == 0 && includeDirect && includeRecip && pmeFamily && !h.timing && h.sortedPmeValid && rank0Share(h) <= 0.1;
    GraphKey key;
    memset(&key, 0, sizeof(key));
    key.pmeEarly = h.pmeEarly;
    key.flags = flags; key.hostPositions = in.hostPositions; key.pbcMode = in.pbcMode;
    for (int k = 0; k < 3; k++) key.cellDim[k] = h.cellDim[k];
    key.numCells = h.numCells;
    key.fusedSort = h.fusedSort; key.beginKernel = h.beginKernel; key.compactSort = h.compactSort;
    key.needEnergy = (flags&SNB_INCLUDE_ENERGY) || !h.derivativeParams.empty();
    key.maxChunks = h.maxChunks;
    key.share = h.comm ? rank0Share(h) : 0.0;
    key.listSkin = h.skinNow;
    key.broadcast = h.comm && h.broadcastPositions;
    key.posqDev = posqDev; key.forcesFixedOut = forcesFixedOut; key.jList = h.jList.p; key.hostPos = hostPos;
    key.posqCorrection = h.ordCorrection; key.atomIndex = h.ordAtomIndex; key.outStride = h.ordStride; key.positionsDouble = h.ordDouble;
    const bool graphable = h.useGraph && !h.timing && (!h.comm || h.balanceFrozen);
    const bool sameKey = h.graphKeyValid && memcmp(&key, &h.graphKey, sizeof(key)) == 0;
    if (!graphable || !sameKey)
        dropGraph(h);

    CUDA_CHECK(cudaEventRecord(h.evBegin, st));
    bool captureFailed = false;
    if (graphable && sameKey) {
        if (!h.graphExec) {
            // second evaluation with this shape: capture it (the first ran eagerly and did every lazy set-up)
            cudaGraph_t graph = nullptr;
            CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                enqueueStep(h, in);
            }
            catch (const FusedSortRefused&) {
                cudaStreamEndCapture(st, &graph);
                if (graph) cudaGraphDestroy(graph);
                graph = nullptr;
                captureFailed = true;       // (the cooperative kernel is the one node kind this graph had not held before)
            }
            catch (...) {
                cudaStreamEndCapture(st, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            if (!captureFailed) {
                CUDA_CHECK(cudaStreamEndCapture(st, &graph));
                cudaError_t err = cudaGraphInstantiate(&h.graphExec, graph, 0);
                cudaGraphDestroy(graph);
                if (err != cudaSuccess) {
                    h.graphExec = nullptr;
                    if (!h.fusedSortUsed)
                        throw SnbException(SNB_ERR_CUDA, std::string("cudaGraphInstantiate: ")+cudaGetErrorString(err));
                    captureFailed = true;
                }
            }
        }
        if (!captureFailed)
            CUDA_CHECK(cudaGraphLaunch(h.graphExec, st));
    }
    else {
        try {
            enqueueStep(h, in);
        }
        catch (const FusedSortRefused&) {
            captureFailed = true;           // the cooperative launch was refused: this evaluation again, with the separate kernels
        }
        h.graphKey = key;
        h.graphKeyValid = graphable && !captureFailed;
    }
    if (captureFailed) {
        // the fused sort kernel could not be launched or captured here: off for this handle, the evaluation runs eagerly
        // with the separate kernels (and is captured the time after next, under its new key)
        cudaGetLastError();
        cudaStreamSynchronize(st);
        cudaGetLastError();
        h.fusedSort = 0;
        h.fusedSortFailed = true;
        dropGraph(h);
        h.graphKeyValid = false;
        CUDA_CHECK(cudaEventRecord(h.evBegin, st));
        enqueueStep(h, in);
    }
    CUDA_CHECK(cudaEventRecord(h.evEnd, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventElapsedTime(&h.lastDeviceMs, h.evBegin, h.evEnd));
    h.listValid = includeDirect;
    if (h.comm && h.rank == 0 && includeDirect && h.sortedAtomPme.p)
        h.sortedPmeValid = true;
    if (h.comm) {
        const NcclApi* nccl = ncclApi(nullptr);
        ncclResult_t asyncErr = ncclSuccess;
        if (nccl && nccl->CommGetAsyncError && nccl->CommGetAsyncError(h.comm, &asyncErr) == ncclSuccess && asyncErr != ncclSuccess)
            throw SnbException(SNB_ERR_CUDA, std::string("NCCL asynchronous error: ")+nccl->GetErrorString(asyncErr));
        for (int k = 0; k < SNB_ENERGY_DOUBLES; k++)
            h.hEnergy.p[k] = (double) h.hEnergyFixed.p[k]*(1.0/SNB_FORCE_SCALE_D);
        // rank 0: its reciprocal-space energies were kept out of the reduction (they were still being computed)
        if (h.rank == 0 && includeRecip)
            for (int k = 0; k < 2*SNB_MAX_SLICES; k++)
                h.hEnergy.p[k] += h.hEnergyPme.p[k];
    }
    if (h.comm && includeDirect && !h.timing && !h.balanceFrozen) {
        // load balance: what was measured one evaluation ago has just come back from the all-reduce, identical on every rank
        const double* got = h.hEnergy.p+32*2*SNB_MAX_SLICES;
        const int W = h.world;
        double D = 0, others = 0;
        for (int r = 0; r < W && r < 8; r++) D += got[2+r];
        for (int r = 1; r < W && r < 8; r++) others += got[10+r]/(W-1);
        const double t0 = got[10];
        if (h.autoBalance && h.rank0Share < 0 && D > 0 && t0 > 0 && W <= 8 && getenv("SNB_RANK0_SHARE") == nullptr) {
            // Newton step on t_others(f0) - t_0(f0) = 0 with dt_0/df0 = D, dt_others/df0 = -D/(W-1); D = summed direct-phase time
            const double cur = rank0Share(h), f0 = cur/(cur+(W-1));
            const double f0new = std::min(1.0/W, std::max(0.0, f0+0.7*(others-t0)*(W-1)/(W*D)));
            h.tunedShare = f0new*(W-1)/(1.0-f0new);
        }
        // this evaluation's own times, for the next all-reduce
        float ms = 0;
        for (int k = 0; k < SNB_ENERGY_FLAGS; k++) h.loadH.p[k] = 0.0;
        if (h.rank < 8 && cudaEventElapsedTime(&ms, h.evDirBegin, h.evDirEnd) == cudaSuccess) h.loadH.p[2+h.rank] = ms;
        if (h.rank < 8 && cudaEventElapsedTime(&ms, h.evBegin, h.evReady) == cudaSuccess) h.loadH.p[10+h.rank] = ms;
        if (h.rank == 0 && (flags&SNB_INCLUDE_RECIPROCAL) && pmeFamily && cudaEventElapsedTime(&ms, h.evRecBegin, h.evRecEnd) == cudaSuccess) h.loadH.p[1] = ms;
        // rank 0 is ready when BOTH its share of direct space and its reciprocal space (own stream, overlapping the reduction) are
        if (h.rank == 0 && (flags&SNB_INCLUDE_RECIPROCAL) && pmeFamily && cudaEventElapsedTime(&ms, h.evBegin, h.evRecEnd) == cudaSuccess)
            h.loadH.p[10] = std::max(h.loadH.p[10], (double) ms);
        h.loadValid = true;
        cudaGetLastError();     // an unrecorded event above must not poison the next evaluation's error check
        if (++h.commEvals >= 8) {
            // the partition has settled: freeze it (a stable launch shape), and let CUDA graphs take over
            h.balanceFrozen = true;
            if (h.rank0Share < 0 && h.tunedShare >= 0) h.rank0Share = h.tunedShare;
        }
    }
    if (includeDirect) {
        const bool anyOverflow = h.comm ? h.hEnergy.p[32*2*SNB_MAX_SLICES] != 0.0 : h.hCounters.p[2] != 0;
        if (h.hCounters.p[2]) {
            // list buffers too small: grow to what the device asked for and repeat
            h.maxChunks = (size_t) (h.hCounters.p[0]*1.25)+64;
            h.jList.alloc(h.maxChunks*32);
            h.chunkOctet.alloc(h.maxChunks);
        }
        if (anyOverflow) {
            h.listKept = false;
            return false;       // all ranks repeat together (the flag travelled with the all-reduce)
        }
    }
    if (h.timing) {
        for (int k = 0; k < 9; k++)
            CUDA_CHECK(cudaEventElapsedTime(&h.phaseMs[k], h.ev[k], h.ev[k+1]));
        CUDA_CHECK(cudaEventElapsedTime(&h.phaseMs[SNB_PHASE_TOTAL], h.ev[0], h.ev[9]));
    }
    if (includeDirect && h.skinNow > 0) {
        // the device has the list of this evaluation's (or an earlier) rebuild: remember what it was built for
        h.listKept = true;
        memcpy(h.listBox, box, sizeof(h.listBox));
        h.listMaxChunks = h.maxChunks;
        h.listShare = h.comm ? rank0Share(h) : 0.0;
        h.listBuiltSkin = h.skinNow;
        h.listEvals++;
        if (h.hCounters.p[SNB_COUNTER_REBUILD]) h.listRebuilds++;
    }
    h.countersOut[0] = h.fusedSortFailed ? -1 : (includeDirect && h.fusedSortUsed ? 1 : 0);
    h.countersOut[1] = h.hCounters.p[0];
    h.countersOut[6] = h.listRebuilds;
    h.countersOut[7] = h.listEvals;
    h.countersOut[3] = h.paramLaunches;
    h.countersOut[4] = h.graphExec ? 1 : 0;
    h.countersOut[5] = includeDirect && h.packedLaunched ? 1 : 0;
    // kernels of THIS library launched by the evaluation (cuFFT's own kernels are not counted)
    long long launches = 0;
    if (includeDirect) launches += (h.fusedSortUsed ? 1 : (h.compactSort && periodic && !(h.skinNow > 0) && h.numCells <= SORT_SMEM_MAX_CELLS) ? 3 : (periodic ? 5 : 6)+(h.numCells > 16384 ? 2 : 0))+1+1+(h.X > 0 ? 1 : 0);    // sort, list, direct, bonded
    if (includeRecip && h.method == SNB_EWALD && h.rank == 0) launches += 2;    // structure factors, forces
    if (includeRecip && pmeFamily && h.rank == 0)      // spread (+ finish or fold), convolution, combine, interpolate
        for (int term = 0; term < (h.method == SNB_LJPME ? 2 : 1); term++)
            launches += h.deterministic ? 5 : 4;
    launches -= h.pmeFoldedLaunches;      // (one or two subsets: no k_pmeCombine launch)
    if (includeForces || h.comm) launches += 1+(h.comm && includeForces ? 1 : 0);
    if (posqDev || h.beginKernelUsed) launches += 1;     // position conversion, or k_begin (which includes it)
    h.countersOut[2] = launches;
    return true;
}

void finishEnergies(snb_handle& h, const double* box, int flags) {
    const bool includeDirect = flags&SNB_INCLUDE_DIRECT, includeRecip = flags&SNB_INCLUDE_RECIPROCAL;
    for (int k = 0; k < 2*h.L; k++) {
        double sum = 0;
        for (int r = 0; r < 32; r++)
            sum += h.hEnergy.p[(size_t) r*2*SNB_MAX_SLICES+k];
        h.sliceEnergies[k] = sum;
    }
    // (multi-GPU: the all-reduce gave every rank the summed slice energies; the host-side terms below are
    // added identically everywhere, so every rank returns the same energies.  Forces land on rank 0.)
    if (includeRecip && (h.method == SNB_EWALD || h.method == SNB_PME || h.method == SNB_LJPME))
        for (int s = 0; s < h.S; s++) {
            h.sliceEnergies[2*(s*(s+3)/2)] += h.hSelf.p[2*s];
            h.sliceEnergies[2*(s*(s+3)/2)+1] += h.hSelf.p[2*s+1];
        }
    if (includeDirect && (h.method == SNB_CUTOFF_PERIODIC || h.method == SNB_EWALD || h.method == SNB_PME)) {
        double volume = box[0]*box[4]*box[8];
        for (int s = 0; s < h.L; s++)
            h.sliceEnergies[2*s+1] += h.dispersionCoef[s]/volume;
    }
    h.lastEnergy = 0;
    if (flags&SNB_INCLUDE_ENERGY)
        for (int s = 0; s < h.L; s++)
            h.lastEnergy += h.lambdaC[s]*h.sliceEnergies[2*s]+h.lambdaV[s]*h.sliceEnergies[2*s+1];
}

void evaluate(snb_handle& h, const double* positions, const void* posqDev, const double* box, const double* globals,
              int flags, void* forcesFixedOut) {
    CUDA_CHECK(cudaSetDevice(h.device));
    const bool periodic = h.method >= SNB_CUTOFF_PERIODIC;
    if (periodic) {
        if (!box)
            throw SnbException(SNB_ERR_INVALID, "periodic method needs box vectors");
        double minAllowedSize = 1.999999*h.cutoff;
        if (box[0] < minAllowedSize || box[4] < minAllowedSize || box[8] < minAllowedSize)
            throw SnbException(SNB_ERR_BOX_TOO_SMALL, "The periodic box size has decreased to less than twice the nonbonded cutoff.");
    }
    std::vector<double> noGlobals(1, 0.0);
    computeParameters(h, globals ? globals : noGlobals.data());
    if (h.N == 0) {
        std::fill(h.sliceEnergies.begin(), h.sliceEnergies.end(), 0.0);
        h.lastEnergy = 0;
        return;
    }
    const double* hostPos = nullptr;
    if (positions && h.comm && h.broadcastPositions && h.rank != 0)
        hostPos = h.hPos.p;         // not read: rank 0's positions arrive by broadcast
    else if (positions) {
        // a page-locked caller buffer is read by the copy engine directly; anything else is staged through the pinned mirror
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, positions) == cudaSuccess && attr.type == cudaMemoryTypeHost)
            hostPos = positions;
        else {
            cudaGetLastError();
            memcpy(h.hPos.p, positions, sizeof(double)*3*(size_t) h.N);
            hostPos = h.hPos.p;
        }
    }
    double unitBox[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double* b = periodic ? box : unitBox;
    for (int attempt = 0; attempt < 4; attempt++)
        if (runOnce(h, b, flags, hostPos, positions ? nullptr : posqDev, forcesFixedOut)) {
            finishEnergies(h, b, flags);
            h.lastFlags = flags;
            return;
        }
    throw SnbException(SNB_ERR_CUDA, "tile list capacity did not converge");
}

int fail(const SnbException& e) {
    g_lastError = e.msg;
    return e.code;
}

}  // namespace

#define SNB_TRY try {
#define SNB_CATCH } catch (const SnbException& e) { return fail(e); } \
    catch (const std::exception& e) { g_lastError = e.what(); return SNB_ERR_INVALID; }

extern "C" {

int snb_create(const snb_desc* desc, snb_handle** out) {
    SNB_TRY
    if (!desc || !out) throw SnbException(SNB_ERR_INVALID, "null argument");
    *out = createHandle(*desc);
    return SNB_OK;
    SNB_CATCH
}

void snb_destroy(snb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    // a captured evaluation holds NCCL operations: the graph must go before its communicator
    if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }
    if (h->comm) {
        if (const NcclApi* nccl = ncclApi(nullptr)) nccl->CommDestroy(h->comm);
        h->comm = nullptr;
    }
    if (h->evCreated)
        for (auto& e : h->ev) cudaEventDestroy(e);
    if (h->graphExec) cudaGraphExecDestroy(h->graphExec);
    for (cudaEvent_t e : {h->evDirBegin, h->evDirEnd, h->evRecBegin, h->evRecEnd, h->evReady})
        if (e) cudaEventDestroy(e);
    if (h->evInputs) cudaEventDestroy(h->evInputs);
    if (h->evBondedDone) cudaEventDestroy(h->evBondedDone);
    if (h->bondedStream) cudaStreamDestroy(h->bondedStream);
    cudaStream_t s = h->stream, s2 = h->pmeStream;
    if (h->evSorted) cudaEventDestroy(h->evSorted);
    if (h->evPmeDone) cudaEventDestroy(h->evPmeDone);
    if (h->evBegin) cudaEventDestroy(h->evBegin);
    if (h->evEnd) cudaEventDestroy(h->evEnd);
    delete h;
    if (s) cudaStreamDestroy(s);
    if (s2) cudaStreamDestroy(s2);
}

int snb_update_parameters(snb_handle* h, const snb_desc* desc) {
    SNB_TRY
    updateHandle(*h, *desc);
    return SNB_OK;
    SNB_CATCH
}

// forces[k] += result[k] on the host (the reference kernel ADDS into the platform's force vector,
// ReferenceOpenMMLabKernels.cpp:35-58).  Memory bound: large systems split the range over a few threads.
static void addForces(double* forces, const double* result, size_t n) {
    auto span = [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; k++)
            forces[k] += result[k];
    };
    const size_t perThread = 1u<<19;      // 4 MiB of doubles per thread at least: below that a thread start costs more than it saves
    unsigned threads = (unsigned) std::min<size_t>(std::min<size_t>(n/perThread, 8), std::max(1u, std::thread::hardware_concurrency()));
    if (threads <= 1) {
        span(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const size_t chunk = (n+threads-1)/threads;
    for (unsigned t = 1; t < threads; t++)
        pool.emplace_back(span, std::min(n, t*chunk), std::min(n, (t+1)*chunk));
    span(0, std::min(n, chunk));
    for (auto& th : pool)
        th.join();
}

int snb_execute(snb_handle* h, const double* positions, const double* box, const double* globals, int flags,
                double* forces, double* energy, double* slice_energies, double* derivatives) {
    SNB_TRY
    if (!positions) throw SnbException(SNB_ERR_INVALID, "positions are null");
    evaluate(*h, positions, nullptr, box, globals, flags, nullptr);
    // (multi-GPU: the reduced forces are delivered on rank 0)
    if ((flags&SNB_INCLUDE_FORCES) && forces && (!h->comm || h->rank == 0))
        addForces(forces, h->hForces.p, 3*(size_t) h->N);
    return snb_read_energies(h, flags, energy, slice_energies, derivatives);
    SNB_CATCH
}

int snb_execute_device(snb_handle* h, const void* posq_dev, const double* box, const double* globals, int flags, void* forces_dev) {
    SNB_TRY
    if (!posq_dev) throw SnbException(SNB_ERR_INVALID, "positions are null");
    evaluate(*h, nullptr, posq_dev, box, globals, flags, forces_dev);
    return SNB_OK;
    SNB_CATCH
}

int snb_execute_device_ordered(snb_handle* h, const snb_device_layout* layout, const double* box, const double* globals, int flags) {
    SNB_TRY
    if (!layout || layout->struct_size != (int32_t) sizeof(snb_device_layout)) throw SnbException(SNB_ERR_INVALID, "snb_device_layout.struct_size does not match this library");
    if (!layout->posq) throw SnbException(SNB_ERR_INVALID, "positions are null");
    if (layout->padded_num_atoms < h->N) throw SnbException(SNB_ERR_INVALID, "padded_num_atoms is smaller than the number of particles");
    struct Scope {      // the layout applies to this evaluation only, whatever way it ends
        snb_handle* h;
        ~Scope() { h->ordCorrection = nullptr; h->ordAtomIndex = nullptr; h->ordStride = 0; h->ordDouble = 0; }
    } scope{h};
    h->ordCorrection = layout->positions_double ? nullptr : layout->posq_correction;
    h->ordAtomIndex = layout->atom_index;
    h->ordStride = layout->padded_num_atoms;
    h->ordDouble = layout->positions_double ? 1 : 0;
    evaluate(*h, nullptr, layout->posq, box, globals, flags, layout->forces);
    return SNB_OK;
    SNB_CATCH
}

int snb_read_energies(snb_handle* h, int flags, double* energy, double* slice_energies, double* derivatives) {
    if (energy) *energy = (flags&SNB_INCLUDE_ENERGY) ? h->lastEnergy : 0.0;
    if (slice_energies)
        std::copy(h->sliceEnergies.begin(), h->sliceEnergies.end(), slice_energies);
    // derivatives: raw slice energies of every (slice, term) bound to the parameter, regardless of
    // includeEnergy (ReferenceOpenMMLabKernels.cpp:252-258)
    if (derivatives)
        for (int s = 0; s < h->L; s++)
            for (int t = 0; t < 2; t++)
                if (h->sliceHasDeriv[s][t])
                    for (size_t k = 0; k < h->derivativeParams.size(); k++)
                        if (h->derivativeParams[k] == h->sliceParam[s][t])
                            derivatives[k] += h->sliceEnergies[2*s+t];
    return SNB_OK;
}

int snb_get_pme_parameters(const snb_handle* h, double* alpha, int* nx, int* ny, int* nz) {
    if (h->method != SNB_PME && h->method != SNB_LJPME) {
        g_lastError = "getPMEParametersInContext: This Context is not using PME or LJPME";
        return SNB_ERR_WRONG_METHOD;
    }
    *alpha = h->alpha; *nx = h->grid[0]; *ny = h->grid[1]; *nz = h->grid[2];
    return SNB_OK;
}

int snb_get_ljpme_parameters(const snb_handle* h, double* alpha, int* nx, int* ny, int* nz) {
    if (h->method != SNB_LJPME) {
        g_lastError = "getPMEParametersInContext: This Context is not using LJPME";
        return SNB_ERR_WRONG_METHOD;
    }
    *alpha = h->dalpha; *nx = h->dgrid[0]; *ny = h->dgrid[1]; *nz = h->dgrid[2];
    return SNB_OK;
}

int snb_dump_pairs(snb_handle* h, int32_t** pairs, int64_t* num_pairs) {
    SNB_TRY
    if (!h->listValid) throw SnbException(SNB_ERR_INVALID, "no tile list has been built (run an execute with direct space first)");
    CUDA_CHECK(cudaSetDevice(h->device));
    DevBuf<unsigned long long> count;
    DevBuf<int2> buf;
    count.alloc(1);
    long long capacity = 0;
    unsigned long long found = 0;
    for (int pass = 0; pass < 2; pass++) {
        CUDA_CHECK(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->counters.p+3, 0, sizeof(int), h->stream));
        DirectArgs da = h->lastDirect;
        buf.alloc(std::max<long long>(capacity, 1));
        da.pairDump = buf.p; da.pairDumpCount = count.p; da.pairDumpCapacity = capacity;
        launchDirect(da, h->stream, h->numSMs);
        CUDA_CHECK(cudaMemcpyAsync(&found, count.p, sizeof(found), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        capacity = (long long) found;
    }
    std::vector<int2> host(std::max<size_t>(found, 1));
    if (found)
        CUDA_CHECK(cudaMemcpy(host.data(), buf.p, sizeof(int2)*found, cudaMemcpyDeviceToHost));
    host.resize(found);
    std::sort(host.begin(), host.end(), [](const int2& a, const int2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    *num_pairs = (int64_t) found;
    *pairs = (int32_t*) malloc(std::max<size_t>(found, 1)*2*sizeof(int32_t));
    for (size_t k = 0; k < found; k++) { (*pairs)[2*k] = host[k].x; (*pairs)[2*k+1] = host[k].y; }
    return SNB_OK;
    SNB_CATCH
}

int snb_count_pairs(snb_handle* h, int64_t* num_pairs) {
    SNB_TRY
    if (!h->listValid) throw SnbException(SNB_ERR_INVALID, "no tile list has been built (run an execute with direct space first)");
    CUDA_CHECK(cudaSetDevice(h->device));
    DevBuf<unsigned long long> count;
    DevBuf<int2> buf;
    count.alloc(1);
    buf.alloc(1);
    unsigned long long found = 0;
    CUDA_CHECK(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), h->stream));
    CUDA_CHECK(cudaMemsetAsync(h->counters.p+3, 0, sizeof(int), h->stream));
    DirectArgs da = h->lastDirect;
    da.pairDump = buf.p; da.pairDumpCount = count.p; da.pairDumpCapacity = 0;
    launchDirect(da, h->stream, h->numSMs);
    CUDA_CHECK(cudaMemcpyAsync(&found, count.p, sizeof(found), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    *num_pairs = (int64_t) found;
    return SNB_OK;
    SNB_CATCH
}

void snb_free(void* p) { free(p); }

/* test hook: capacity of the tile list in chunks (the next evaluation re-allocates; too small a value exercises the
 * overflow -> grow -> repeat path) */
int snb_set_list_capacity(snb_handle* h, int64_t chunks) {
    SNB_TRY
    if (!h || chunks < 1) throw SnbException(SNB_ERR_INVALID, "bad list capacity");
    CUDA_CHECK(cudaSetDevice(h->device));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    dropGraph(*h);
    h->listKept = false;
    h->jList.release(); h->chunkOctet.release();
    h->maxChunks = (size_t) chunks;
    h->jList.alloc(h->maxChunks*32);
    h->chunkOctet.alloc(h->maxChunks);
    return SNB_OK;
    SNB_CATCH
}

/* Tile-list reuse: skin_nm > 0 builds the list with cutoff + skin and keeps it (and the sorted order) until an atom has moved
 * more than skin/2, the box changes, or the buffers / the multi-GPU partition do; 0 (default) rebuilds every evaluation. */
int snb_set_list_skin(snb_handle* h, double skin_nm) {
    SNB_TRY
    if (!h || !(skin_nm >= 0.0) || skin_nm > 10.0) throw SnbException(SNB_ERR_INVALID, "bad list skin");
    CUDA_CHECK(cudaSetDevice(h->device));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (skin_nm > 0 && h->N > 0 && !h->posRef.p) {
        h->posRef.alloc(3*(size_t) h->N);
        h->wrapImage.alloc(h->N);
        h->listState.alloc(4);
        CUDA_CHECK(cudaMemset(h->listState.p, 0, 4*sizeof(int)));
    }
    h->listSkin = skin_nm;
    h->listKept = false;
    return SNB_OK;
    SNB_CATCH
}

int snb_set_timing(snb_handle* h, int enabled) { h->timing = enabled != 0; return SNB_OK; }

int snb_get_phase_times(snb_handle* h, float* ms) {
    for (int k = 0; k < SNB_NUM_PHASES; k++) ms[k] = h->phaseMs[k];
    return SNB_OK;
}

int snb_get_last_device_ms(snb_handle* h, float* ms) { *ms = h->lastDeviceMs; return SNB_OK; }

int snb_get_counters(snb_handle* h, int64_t* out) {
    for (int k = 0; k < 8; k++) out[k] = h->countersOut[k];
    return SNB_OK;
}

int snb_nccl_unique_id(void* id128) {
    SNB_TRY
    const char* why = nullptr;
    const NcclApi* nccl = ncclApi(&why);
    if (!nccl) throw SnbException(SNB_ERR_UNSUPPORTED, std::string("NCCL is not available: ")+why);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NCCL_CHECK(nccl, nccl->GetUniqueId((ncclUniqueId*) id128));
    return SNB_OK;
    SNB_CATCH
}

int snb_comm_init(snb_handle* h, const void* id128, int rank, int world_size) {
    SNB_TRY
    if (!h || !id128 || world_size < 1 || rank < 0 || rank >= world_size)
        throw SnbException(SNB_ERR_INVALID, "bad communicator arguments");
    const char* why = nullptr;
    const NcclApi* nccl = ncclApi(&why);
    if (!nccl) throw SnbException(SNB_ERR_UNSUPPORTED, std::string("NCCL is not available: ")+why);
    CUDA_CHECK(cudaSetDevice(h->device));
    dropGraph(*h);
    if (h->comm) { nccl->CommDestroy(h->comm); h->comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCCL_CHECK(nccl, nccl->CommInitRank(&h->comm, world_size, id, rank));
    h->rank = rank;
    h->world = world_size;
    h->reduceSend.alloc(3*(size_t) h->N+SNB_ENERGY_DOUBLES);
    h->reduceRecv.alloc(3*(size_t) h->N+SNB_ENERGY_DOUBLES);
    h->loadH.alloc(SNB_ENERGY_FLAGS);
    if (!h->evDirBegin) {
        CUDA_CHECK(cudaEventCreate(&h->evDirBegin)); CUDA_CHECK(cudaEventCreate(&h->evDirEnd));
        CUDA_CHECK(cudaEventCreate(&h->evRecBegin)); CUDA_CHECK(cudaEventCreate(&h->evRecEnd));
        CUDA_CHECK(cudaEventCreate(&h->evReady));
    }
    h->loadValid = false; h->tunedShare = -1;
    h->hEnergyFixed.alloc(SNB_ENERGY_DOUBLES);
    if (rank == 0) {
        h->forcePme.alloc(3*(size_t) h->NP);
        h->energyPme.alloc(2*SNB_MAX_SLICES);
        h->hEnergyPme.alloc(2*SNB_MAX_SLICES);
        std::fill(h->hEnergyPme.p, h->hEnergyPme.p+2*SNB_MAX_SLICES, 0.0);
        if (world_size >= 3 && getenv("SNB_NO_EARLY_PME") == nullptr)      // (two ranks: rank 0 always keeps a share of the list)
            h->sortedAtomPme.alloc(h->NP);
    }
    h->sortedPmeValid = false;
    h->listKept = false;
    h->balanceFrozen = false; h->commEvals = 0;
    return SNB_OK;
    SNB_CATCH
}

int snb_set_rank0_share(snb_handle* h, double share) { h->rank0Share = share; return SNB_OK; }

int snb_set_comm_options(snb_handle* h, int broadcast_positions) { h->broadcastPositions = broadcast_positions != 0; return SNB_OK; }

int snb_host_parameters(const snb_desc* desc, double* alpha, int32_t grid[3], double* dispersion_alpha, int32_t dispersion_grid[3],
                        int32_t kmax[3], double* dispersion) {
    SNB_TRY
    if (!desc) throw SnbException(SNB_ERR_INVALID, "descriptor is null");
    const snb_desc& d = *desc;
    validateDesc(d);
    double a = 0.0, da = 0.0;
    int g[3] = {0, 0, 0}, dg[3] = {0, 0, 0}, km[3] = {0, 0, 0};
    if (d.method == SNB_PME || d.method == SNB_LJPME)
        calcPMEParameters(d, false, a, g);
    if (d.method == SNB_LJPME)
        calcPMEParameters(d, true, da, dg);
    if (d.method == SNB_EWALD)
        calcEwaldParameters(d, a, km);
    if (alpha) *alpha = a;
    if (dispersion_alpha) *dispersion_alpha = da;
    for (int k = 0; k < 3; k++) {
        if (grid) grid[k] = g[k];
        if (dispersion_grid) dispersion_grid[k] = dg[k];
        if (kmax) kmax[k] = km[k];
    }
    if (dispersion) {
        const int L = d.num_subsets*(d.num_subsets+1)/2;
        const std::vector<double> coef = d.use_dispersion_correction ? calcDispersionCorrections(d) : std::vector<double>(L, 0.0);
        std::copy(coef.begin(), coef.end(), dispersion);
    }
    return SNB_OK;
    SNB_CATCH
}

int snb_partition(int64_t n, int world_size, int rank, double rank0_share, int64_t* lo, int64_t* hi) {
    if (!lo || !hi || world_size < 1 || rank < 0 || rank >= world_size || n < 0) {
        g_lastError = "bad partition arguments";
        return SNB_ERR_INVALID;
    }
    partitionRange(n, world_size, rank, rank0_share, lo, hi);
    return SNB_OK;
}

const char* snb_last_error(void) { return g_lastError.c_str(); }
const char* snb_version(void) { return "slicednb-b200 0.1 (sm_100a)"; }

}  // extern "C"
