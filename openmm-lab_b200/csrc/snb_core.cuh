// Internal declarations of the B200 SlicedNonbondedForce core (sm_100a).
//
// Layout in HBM (per handle), N atoms padded to NP = 32*numBlocks:
//   posd      double[N][3]   positions as given (original order) -- exact cutoff predicate, PME, bonded
//   params4   float4[N]      (q, sigma/2, 2 sqrt(eps), bitcast subset)         original order
//   paramsD   double[N][3]   (sigma/2, 2 sqrt(eps), q) in double               original order (bonded kernel)
//   sortedAtom int[NP]       original index of sorted slot k (-1 = padding)
//   sortedPos  int[N]        inverse permutation
//   posqLocal float4[NP]     (x,y,z relative to the block centre, q)           sorted order
//   sparams   float4[NP]     (sigma/2, 2 sqrt(eps), bitcast subset, c6)        sorted order
//   blockCenter float4[NB] (2^-10 nm lattice: exact in float), blockExt float4[NB] (half extents)
//   jList     int2[cap]      (sorted j, exclusion mask over the 32 i lanes), 32 entries per chunk
//   chunkOctet int[cap]      octet (8 i atoms) of every chunk; bit 30: per-pair periodic wrapping
//   forceSorted / forceOrig  long long[3][NP]  2^32 fixed point
//   energyBuf double[numSlices][2] (+ counters)
#ifndef SNB_CORE_CUH_
#define SNB_CORE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slicednb.h"
#include "../../include/snb_constants.h"

#define SNB_MAX_SUBSETS 8
#define SNB_MAX_SLICES (SNB_MAX_SUBSETS*(SNB_MAX_SUBSETS+1)/2)
#define SNB_BLOCK 32
#define SNB_FORCE_SCALE 4294967296.0f       /* 2^32, as the reference's fixed-point forces (realToFixedPoint.cu:1-2) */
#define SNB_FORCE_SCALE_D 4294967296.0
#define SNB_NARROW_MAX_EXTENT 1.0f         /* nm: largest block half extent for which the narrow cutoff band is valid */
#define SNB_COUNTER_REBUILD 6               /* counters[]: nonzero = this evaluation rebuilds the sorted order and the tile list (list reuse on) */
#define SNB_COUNTER_BARRIER 7               /* counters[]: arrive count of the fused sort kernel's grid-wide barriers */
#define SNB_COUNTER_MAX_EXTENT 5            /* counters[]: float bits of the largest block half extent of this evaluation */

enum { PBC_NONE = 0, PBC_SHIFT = 1, PBC_PAIR_ORTHO = 2, PBC_PAIR_TRICLINIC = 3 };
enum { KIND_PLAIN = 0, KIND_RF = 1, KIND_EWALD = 2, KIND_LJPME = 3 };

struct BoxF {
    float ax, bx, by, cx, cy, cz;       // a=(ax,0,0) b=(bx,by,0) c=(cx,cy,cz)
    float invAx, invBy, invCz;
};
struct BoxD {
    double ax, bx, by, cx, cy, cz;
    double recip[3][3];                 // lower triangular, frac = pos * recip (ReferencePME.cpp:186-194)
};

struct DirectParams {
    int numAtoms, numBlocks, pbcMode, periodic;
    // |r2 - cutoff2| < band -> the exact double predicate decides.  Narrow band: j atoms brought into the octet's frame
    // with the periodic image taken in double (fix-up pass), every coordinate < ~2 nm: |error(r2)| < 5.2e-7 rc + 1.5e-7 rc^2.
    // Wide band: per-pair wrapping in single precision (triclinic boxes, octets wider than halfBox - cutoff) or blocks
    // wider than SNB_NARROW_MAX_EXTENT, where the error grows with the box / block size.
    float bandTol, bandTolWide;
    double cutoff2d;
    double alphad, krfd, crfd, switchDistd, invSwitchWidthd;
    double dalphad, dispShiftExpd, invCut6d;   // LJPME: alpha_d, [1-exp(-d_c^2)(1+d_c^2+d_c^4/2)], rc^-6
};

// The pair arithmetic's constants, precomputed on the host in the kernel's own precision: they sit in the
// kernel-parameter constant bank and are used as instruction operands (no in-loop conversions).
template <typename R> struct DirectConsts {
    R cutoffHi, cutoffLo;               // cutoff^2 +- band: inside the band the exact double predicate decides
    R cutoffHiWide, cutoffLoWide;       // the wider band of chunks whose r^2 carries box-sized rounding errors (see bandTol)
    R alpha, alpha2, halfAlpha, negAlpha2Log2e, twoAlphaOverSqrtPi;
    R krf, crf, twoKrf;
    R switchDist, invSwitchWidth;
    R dalpha2, dispShiftExp, invCut6;
};

// what the exact (double) cutoff predicate reads; lives in device memory so that the rare out-of-line
// call takes a pointer instead of forcing the kernel's parameter block onto the stack
struct ExactParams {
    BoxD boxd;
    double cutoff2d;
    int periodic;
};

// Everything that may change from one evaluation to the next WITHOUT changing a launch shape: the box
// and the slice lambdas.  One copy in device memory, refreshed by a single host->device copy from a
// pinned mirror at the head of every evaluation, so that the whole evaluation can be replayed as a
// CUDA graph (api.cu) with no per-node parameter updates.
struct StepParams {
    ExactParams exact;                  // exact.boxd is THE double-precision box
    BoxF box;
    float shiftLimit[3];                // halfBox - cutoff per axis (one-image-per-j-atom shortcut)
    double lambdaC[SNB_MAX_SLICES], lambdaV[SNB_MAX_SLICES];
    // tile-list reuse (ListSkin > 0): squared displacement since the last build beyond which the list is rebuilt, and the
    // host's own reasons for a rebuild (first evaluation, box / capacity / partition / parameters changed)
    double skinHalf2;
    int forceRebuild;
};

// ---- small device helpers ---------------------------------------------------------
__device__ __forceinline__ int sliceOf(int a, int b) { return a > b ? a*(a+1)/2+b : b*(b+1)/2+a; }

__device__ __forceinline__ float fastExp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fastRsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fastRcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// erfc(x) = exp(-x^2) * P(u), u = (t - m)/h, t = 1/(1 + x/2); degree-9 Chebyshev fit of erfcx on
// x in [0,4]: fit error 6e-10, float evaluation error < 5e-7 relative (the reference's single-precision
// kernel uses an Abramowitz-Stegun form with 1.5e-7 ABSOLUTE error, coulombLennardJones.cc:18-23,
// which is ~1e-3 relative at the cutoff and cannot meet the 1e-5 parity bar).
__device__ __forceinline__ float erfcxPoly(float x) {
    const float m = 0.66666666666666663f, h = 0.33333333333333337f;   // t in [1/3, 1]
    float t = fastRcp(fmaf(0.5f, x, 1.0f));
    float u = (t-m)*(1.0f/h);
    float p = -7.5385177448e-07f;
    p = fmaf(p, u, -8.2052179778e-06f);
    p = fmaf(p, u, 3.4861879002e-05f);
    p = fmaf(p, u, 9.8919924926e-05f);
    p = fmaf(p, u, -8.6138957608e-04f);
    p = fmaf(p, u, -1.6015017984e-03f);
    p = fmaf(p, u, 2.2509531568e-02f);
    p = fmaf(p, u, 1.4242693719e-01f);
    p = fmaf(p, u, 4.0981802125e-01f);
    p = fmaf(p, u, 4.2758357745e-01f);
    return p;
}

// same polynomial, taking t = 1/(1 + x/2) (the caller folds alpha into the reciprocal's argument)
__device__ __forceinline__ float erfcxPolyT(float t) {
    const float u = fmaf(t, 3.0f, -2.0f);
    float p = -7.5385177448e-07f;
    p = fmaf(p, u, -8.2052179778e-06f);
    p = fmaf(p, u, 3.4861879002e-05f);
    p = fmaf(p, u, 9.8919924926e-05f);
    p = fmaf(p, u, -8.6138957608e-04f);
    p = fmaf(p, u, -1.6015017984e-03f);
    p = fmaf(p, u, 2.2509531568e-02f);
    p = fmaf(p, u, 1.4242693719e-01f);
    p = fmaf(p, u, 4.0981802125e-01f);
    p = fmaf(p, u, 4.2758357745e-01f);
    return p;
}

__device__ __forceinline__ long long toFixed(float f) { return __float2ll_rn(f*SNB_FORCE_SCALE); }
__device__ __forceinline__ long long toFixedD(double f) { return __double2ll_rn(f*SNB_FORCE_SCALE_D); }
__device__ __forceinline__ void addFixed(long long* p, long long v) {
    atomicAdd((unsigned long long*) p, (unsigned long long) v);
}

// OpenMM-style periodic displacement in double (ReferenceForce::getDeltaRPeriodic semantics),
// evaluated without FMA contraction so it matches the CPU oracle bit for bit.
__device__ __forceinline__ void deltaExact(const double* pi, const double* pj, const BoxD& b, bool periodic, double d[3]) {
    d[0] = __dsub_rn(pj[0], pi[0]); d[1] = __dsub_rn(pj[1], pi[1]); d[2] = __dsub_rn(pj[2], pi[2]);
    if (periodic) {
        double s = floor(__dadd_rn(__ddiv_rn(d[2], b.cz), 0.5));
        d[0] = __dsub_rn(d[0], __dmul_rn(b.cx, s)); d[1] = __dsub_rn(d[1], __dmul_rn(b.cy, s)); d[2] = __dsub_rn(d[2], __dmul_rn(b.cz, s));
        s = floor(__dadd_rn(__ddiv_rn(d[1], b.by), 0.5));
        d[0] = __dsub_rn(d[0], __dmul_rn(b.bx, s)); d[1] = __dsub_rn(d[1], __dmul_rn(b.by, s));
        s = floor(__dadd_rn(__ddiv_rn(d[0], b.ax), 0.5));
        d[0] = __dsub_rn(d[0], __dmul_rn(b.ax, s));
    }
}
__device__ __forceinline__ double dot3Exact(const double d[3]) {
    return __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
}

#endif
