// Argument blocks and launchers of the device stages (one .cu per stage).
#ifndef SNB_KERNELS_CUH_
#define SNB_KERNELS_CUH_

#include "snb_core.cuh"

#include <cuda.h>      /* CUtensorMap (type only: the encoder is resolved at run time through cudaGetDriverEntryPoint) */

// ---- computeParameters on the device (params.cu)
#define SNB_PARAM_MAX_BLOCKS 296
struct ParamArgs {
    int numAtoms, numExceptions;
    double selfScaleC;              // ONE_4PI_EPS0 alpha / sqrt(pi)   (0 outside the Ewald family)
    double selfScaleV;              // alpha_d^6 64 / 12               (0 unless LJPME)
    const double* globals;          // [G] current values of the Context parameters
    const double* baseParticle;     // [N][3] charge, sigma, epsilon as given
    const int* particleOffsetStart; // [N+1] CSR over particles, or null when the force has no particle offsets
    const double4* particleOffsets; // (chargeScale, sigmaScale, epsilonScale, parameter index)
    const double* baseException;    // [X][3] chargeProd, sigma, epsilon as given
    const int* exceptionOffsetStart;
    const double4* exceptionOffsets;
    const int* subset;              // [N]
    double* paramsD;                // out [N][3] sigma/2, 2 sqrt(eps), q
    float4* params4;                // out [N]    q, sigma/2, 2 sqrt(eps), subset bits
    double* exception14;            // out [X][3] sigma, 4 eps, chargeProd
    double* selfPartial;            // scratch [SNB_PARAM_MAX_BLOCKS][2*SNB_MAX_SUBSETS]
    double* selfOut;                // out [SNB_MAX_SUBSETS][2] self energies per subset (Coulomb, dispersion)
    unsigned* ticket;               // zero between launches
};
void launchComputeParameters(const ParamArgs& a, cudaStream_t stream);

struct SortArgs {
    int numAtoms, paddedAtoms, numBlocks, numCells, periodic;
    int cellDim[3];
    int* counters;
    float sqrtK;            // sqrt(ONE_4PI_EPS0): direct-space charges carry it
    const StepParams* sp;   // box
    const double* posd;
    const float4* params4;
    const double* paramsD;
    const int* cellRank;
    double* extent;
    int *cellCount, *cellStart, *atomCell, *atomSlot, *cellScratch, *sortedAtom, *sortedPos;
    int* tileSum;           // [ceil(numCells/1024)] scratch of the tiled scan (large cell grids)
    int2* cellRange;        // [numCells] sorted range (first, end) of every cell, by LINEAR cell index
    float4 *posqLocal, *sparams;
    double4 *posqLocalD, *sparamsD;  // double-precision mode only (else null)
    float4* blockCenter;    // [numBlocks] centre of every block (exact in float, see k_blockBounds)
    float4* blockExt;
    float4 *octetCenter, *octetExt;  // [4*numBlocks] block-local centre / half extents of every octet
    // tile-list reuse (off: reuse = 0 and the pointers are null).  The sorted order and the list are rebuilt only when
    // k_listCheck raises counters[SNB_COUNTER_REBUILD]; k_blockBounds runs every evaluation on the kept order.
    int reuse;
    double* posRef;         // [N][3] positions at the last rebuild
    float4* wrapImage;      // [N] periodic image (whole box vectors) every atom was wrapped by at the last rebuild
    // the sort as one cooperative kernel (k_sortFused: periodic, no list reuse, at most SORT_FUSED_MAX_CELLS cells)
    int fused, numSMs;
    int compact;            // three kernels instead of five where it applies (periodic, no list reuse, <= SORT_SMEM_MAX_CELLS cells)
    int* slotCell;          // [N] cell (rank along the curve) of every slot of the scattered order
};
#define SORT_SMEM_MAX_CELLS 8192        /* the scan of the cell counts fits every CTA's shared memory */
#define SORT_FUSED_MAX_CELLS 4096       /* measured on B200: 648 / 3 000 atoms +5 % / +4 % evaluations/s, 23 558 atoms (5 832 cells) -1.5 % */
// usedFused (optional): whether the fused kernel took the launch; the error is the cooperative launch's (the caller may
// fall back to the separate kernels by clearing `fused`)
cudaError_t launchSort(const SortArgs& a, cudaStream_t stream, bool* usedFused = nullptr);
int sortFusedGrid(const SortArgs& a);     // > 0: launchSort will take the fused kernel with this grid

struct ListArgs {
    int numAtoms, numBlocks, pbcMode, periodic, useCutoff;
    float cutoff;           // padded by a safety margin for the single-precision tests
    const StepParams* sp;   // box, shiftLimit
    const int* sortedAtom;
    const int* sortedPos;
    const float4* posqLocal;
    const float4* blockCenter;
    const float4* blockExt;
    const float4 *octetCenter, *octetExt;
    const int* exclStart;   // CSR over ORIGINAL atom indices: partners of each atom (both directions)
    const int* exclList;
    int2* jList;
    int* chunkOctet;        // [maxChunks] octet of every chunk (bit 30: per-pair periodic wrapping)
    int* counters;          // [0] chunks allocated, [2] overflow flag
    int maxChunks;
    const int2* cellRange;  // [numCells] sorted range of every cell, by linear cell index ((x*dimY+y)*dimZ+z)
    int cellDim[3];
    const double* extent;   // non-periodic methods: bounding box of all atoms (min[3], max[3])
    int blockLo, blockHi;   // this rank's i blocks
    int bitmapWords;        // ceil(numBlocks/32)
    int refine;             // test the candidates against the octet's atoms, not only their box (tuning switch, default on)
    int reuse;              // list reuse on: build only when counters[SNB_COUNTER_REBUILD] is raised
    float skin;             // (cutoff already includes it) octets this close to the one-image limit wrap per pair: they may grow
};
cudaError_t launchBuildList(const ListArgs& a, cudaStream_t stream, int numSMs);

#ifndef DIRECT_CTAS_PER_SM
#define DIRECT_CTAS_PER_SM 16       /* one-warp CTAs of the direct-space kernel resident per SM (register budget 65536/32/this) */
#endif

#ifndef DIRECT_WIDE_CTAS_PER_SM
#define DIRECT_WIDE_CTAS_PER_SM 12  /* the same for more than two subsets (per-subset lambda / energy tables: 168 registers) */
#endif

#ifndef DIRECT_UNROLL
#define DIRECT_UNROLL 8
#endif

struct DirectArgs {
    DirectParams p;
    DirectConsts<float> cf;
    DirectConsts<double> cd;
    const StepParams* sp;          // box, lambdas, exact-predicate parameters
    int kind, useSwitch, numSubsets, needEnergy, doublePrecision;
    int usePacked;                 // the packed-FP32 kernel takes the launch where it applies (direct_packed.cu)
    const int* sortedAtom;
    const float4* posqLocal;
    const float4* sparams;
    const double4* posqLocalD;     // double-precision mode only
    const double4* sparamsD;
    const float4* blockCenter;
    const float4* octetCenter;
    const double* posd;
    const int2* jList;
    const int* chunkOctet;
    int maxChunks;
    int* counters;
    int* listState;             // list reuse: [0] chunks of the kept list (device-resident between evaluations), or null
    long long* forceSorted;     // [3][paddedAtoms]
    int paddedAtoms;
    double* energyBuf;          // [numSlices][2]
    int2* pairDump;             // optional (tests): original-index pairs
    unsigned long long* pairDumpCount;
    long long pairDumpCapacity;
};
bool launchDirect(const DirectArgs& a, cudaStream_t stream, int numSMs);      // true: the packed-FP32 kernel ran

struct BondedArgs {
    int numExceptions, periodicExceptions, ewald, ljpme, includeForces;
    int firstException;             // this rank's contiguous range is [firstException, firstException+numExceptions)
    double alpha, dalpha;
    const StepParams* sp;           // box, lambdas
    const double* posd;
    const double* paramsD;          // [N][3] sigma/2, 2 sqrt(eps), q
    const int* subset;
    const int2* exceptionAtoms;     // [X]
    const double* exception14;      // [X][3] sigma, 4 eps, chargeProd (after offsets)
    const int* exceptionIs14;       // [X]
    long long* forceOrig;           // [3][paddedAtoms]
    int paddedAtoms;
    double* energyBuf;
};
void launchBonded(const BondedArgs& a, cudaStream_t stream);

struct PmeArgs {
    int numAtoms, numSubsets, term;  // term 0 Coulomb, 1 dispersion
    int nx, ny, nz, doublePrecision;
    double alpha;
    const StepParams* sp;           // box, lambdas (term 0: lambdaC, term 1: lambdaV)
    const double* posd;
    const int* sortedAtom;
    const float4* params4;          // q in .x ; sigma/2, 2 sqrt(eps) in .y .z (dispersion "charge" = 8 s^3 e) ; subset bits in .w
    const double* paramsD;          // [N][3] sigma/2, 2 sqrt(eps), q
    void* grid;                     // [S][nx][ny][nz] real (float or double)
    float* gridPad;                 // brick path (pme_brick.cu): [S][nx+4][ny+4][nzp] padded copy the TMA bricks address, or null
    int nzp;                        // its row length: nz+4 rounded up to a multiple of 4 floats
    long long* gridFixed;           // same shape, 2^40 fixed point (deterministic spreading), or null
    void* cgrid;                    // [S][nx][ny][nz/2+1] complex
    void* eterm;                    // [nx][ny][nz/2+1]
    const double* moduli[3];
    long long* forceOrig;
    int paddedAtoms;
    double* energyBuf;
};
void launchPmeEterm(const PmeArgs& a, cudaStream_t stream);
void launchPmeZero(const PmeArgs& a, cudaStream_t stream);
void launchPmeSpread(const PmeArgs& a, cudaStream_t stream, bool zeroed = false);
void launchPmeConvolution(const PmeArgs& a, cudaStream_t stream);
bool pmeInterpolateFolds(const PmeArgs& a);   // the per-atom interpolation does the lambda combination itself (no k_pmeCombine launch)
void launchPmeInterpolate(const PmeArgs& a, cudaStream_t stream);
// brick path (mixed precision, sorted atoms): shared-memory grid bricks moved by TMA
bool makeBrickTensorMap(CUtensorMap* map, float* gridPad, int numAtoms, int numSubsets, int nx, int ny, int nzp);
cudaError_t launchPmeInterpolateBrick(const PmeArgs& a, const CUtensorMap& map, cudaStream_t stream);

struct EwaldArgs {
    int numAtoms, numSubsets, numK, paddedAtoms;
    double alpha;
    const StepParams* sp;           // box, Coulomb lambdas
    const double* posd;
    const double* paramsD;          // [N][3] sigma/2, 2 sqrt(eps), q
    const int* subset;
    const int3* kvec;               // [numK] the half space of reciprocal vectors, in units of 2 pi / box length
    double2* structure;             // [numK][numSubsets] structure factors
    long long* forceOrig;
    double* energyBuf;              // or null: no energies wanted
};
void launchEwaldReciprocal(const EwaldArgs& a, bool includeForces, cudaStream_t stream);

struct FinishArgs {
    int numAtoms, paddedAtoms;
    const int* sortedPos;
    const long long* forceSorted;
    const long long* forceOrig;
    double* forcesOut;              // [N][3] host-visible staging (device), or null
    long long* forcesFixedOut;      // [3][outStride] added to, or null
    const int* outIndex;            // null: forcesFixedOut is in original atom order; else slot t holds atom outIndex[t] (the caller's order)
    int outStride;                  // component stride of forcesFixedOut (numAtoms, or the caller's padded atom count)
    long long* reduceOut;           // [3][N] overwritten: this rank's contribution to the NCCL force reduction, or null
    const int* counters;            // list-build counters; [2] = overflow flag
    const double* energyIn;         // multi-GPU: the slice-energy table + flags, appended to reduceOut in 2^32 fixed point
    int numEnergy, overflowSlot;    // entries of energyIn; which of them takes counters[2] instead
    // single GPU: the kernel itself stores the energy table and the eight list counters into the host's pinned mirrors
    // (mapped memory; 2.5 KB of posted writes) instead of two device->host copies behind it
    double* hostEnergy;             // device-side address of the pinned energy mirror, or null
    const double* energySrc;
    int numEnergyHost;
    int* hostCounters;
};
// after the force reduction (root only): reduced [3][N] fixed-point forces -> the caller's layout
// overflow: the reduced list-overflow flag (device pointer, nonzero = the pass is being discarded) or null
// extra: forces to add on top of the reduced ones ([3][paddedAtoms], original order), or null
void launchEmitReduced(const long long* reduced, int numAtoms, int paddedAtoms, const long long* extra, double* forcesOut, long long* forcesFixedOut,
                       const int* outIndex, int outStride, const long long* overflow, cudaStream_t stream);
void launchFinish(const FinishArgs& a, cudaStream_t stream);
// device positions -> the double array in original atom order.  posq: float4 (or double4 when positionsDouble) per slot;
// correction: float4 per slot added to a float4 position (OpenMM's mixed precision), or null; atomIndex: original index of
// the atom in every slot, or null when the slots are in original order
// First kernel of a single-GPU evaluation: the step's parameters from the host's pinned mirror (mapped memory) to their device
// copy, every accumulator of the evaluation zeroed, and the positions of a device-resident caller converted -- one launch
// for what used to be a copy, a memset and a kernel at the head of every chain of the evaluation.
struct BeginArgs {
    const int* stepHost;            // device-side address of the pinned StepParams
    int* stepDev;
    int stepWords;                  // sizeof(StepParams)/4
    unsigned long long* zero;       // the block of accumulators (forces, energies, counters)
    size_t zeroWords;               // its size in 8-byte words
    const void* posq;               // device-resident positions (null: the host's are already in posd)
    const float4* correction;
    const int* atomIndex;
    int positionsDouble;
    double* posd;
    int numAtoms;
};
void launchBegin(const BeginArgs& a, cudaStream_t stream);
void launchConvertPositions(const void* posq, const void* correction, const int* atomIndex, int positionsDouble, double* posd, int numAtoms, cudaStream_t stream);

#endif
