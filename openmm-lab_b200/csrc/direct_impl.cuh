// Direct-space tile kernel (north-star subsystem 2).
//
// One warp per work item = one octet of i atoms against up to SNB_ITEM_CHUNKS chunks of 32 compacted
// j atoms.  The octet is held four times over in the warp (lane = 8*copy + k, atom k of the octet);
// copy c owns j slots 8c..8c+7 of the chunk, so a chunk is 8 steps of 32 pairs.  A chunk's j atoms are
// gathered with coalesced float4 loads, shifted into the octet's frame (one periodic image per j atom
// where the box allows it) and staged in shared memory.  In step s lane (c,k) works on slot
// 8c + (k+s)%8: every step touches 32 different pairs, the j-force accumulators travel between the 8
// lanes of a copy by warp shuffle and are home after 8 steps, so no reduction tree is needed; the four
// copies' i forces are summed by shuffle once per work item.
// Pair arithmetic is FP32 (MUFU rsqrt + Newton step, MUFU ex2, polynomial erfc); i forces are carried
// in FP64 across chunks; all forces leave as 2^32 fixed-point 64-bit atomics (order independent =>
// bit-reproducible); slice energies are accumulated per j subset in FP32 per chunk and FP64 beyond.
//
// Pair arithmetic follows the reference's Reference platform:
//   Ewald/PME/LJPME direct:  platforms/reference/src/ReferenceSlicedLJCoulombIxn.cpp:351-429
//   cutoff / no cutoff:      ReferenceSlicedLJCoulombIxn.cpp:553-611 (reaction field :64-65)
// (behavioural cross-reference on the GPU side: platforms/common/src/kernels/coulombLennardJones.cc:1-124).
#include "kernels.cuh"

#define DIRECT_CTAS_PER_SM 24
#define ENERGY_REPLICAS 32      /* copies of the slice-energy table, to spread the final atomics */

template <int KIND> void launchDirectKind(const DirectArgs& a, cudaStream_t st, int grid);

// Precision traits: R = float is the mixed-precision path; R = double is the double-precision mode
// (the analogue of the reference CUDA platform's Precision=double), used to prove parity to 1e-5 on
// every quantity.
template <typename R> struct Real;
template <> struct Real<float> {
    typedef float4 vec4;
    static __device__ __forceinline__ const float4* pos(const DirectArgs& a) { return a.posqLocal; }
    static __device__ __forceinline__ const float4* prm(const DirectArgs& a) { return a.sparams; }
    static __device__ __forceinline__ const DirectConsts<float>& consts(const DirectArgs& a) { return a.cf; }
    static __device__ __forceinline__ float rsqrt(float r2) {
        // MUFU.RSQ (2 ulp) + one Newton step: the LJ force goes as invR^14, so the seed's error
        // would otherwise dominate the pair error
        float y = fastRsqrt(r2);
        return y*fmaf(-0.5f*r2, y*y, 1.5f);
    }
    static __device__ __forceinline__ float expNeg(float x) { return fastExp2(-1.4426950408889634f*x); }    // exp(-x)
    // ex = exp(-alpha^2 r^2), erfcAR = erfc(alpha r): one MUFU.EX2, one MUFU.RCP, ten FMAs
    static __device__ __forceinline__ void erfcTerms(const DirectConsts<float>& c, float r, float r2, float& ex, float& erfcAR) {
        ex = fastExp2(r2*c.negAlpha2Log2e);
        erfcAR = ex*erfcxPolyT(fastRcp(fmaf(r, c.halfAlpha, 1.0f)));
    }
    static __device__ __forceinline__ float rint(float x) { return rintf(x); }
    static __device__ __forceinline__ long long fixed(float f) { return toFixed(f); }
};
template <> struct Real<double> {
    typedef double4 vec4;
    static __device__ __forceinline__ const double4* pos(const DirectArgs& a) { return a.posqLocalD; }
    static __device__ __forceinline__ const double4* prm(const DirectArgs& a) { return a.sparamsD; }
    static __device__ __forceinline__ const DirectConsts<double>& consts(const DirectArgs& a) { return a.cd; }
    static __device__ __forceinline__ double rsqrt(double r2) { return 1.0/sqrt(r2); }
    static __device__ __forceinline__ double expNeg(double x) { return exp(-x); }
    static __device__ __forceinline__ void erfcTerms(const DirectConsts<double>& c, double r, double r2, double& ex, double& erfcAR) {
        ex = exp(-c.alpha2*r2);
        erfcAR = erfc(c.alpha*r);
    }
    static __device__ __forceinline__ double rint(double x) { return ::rint(x); }
    static __device__ __forceinline__ long long fixed(double f) { return toFixedD(f); }
};

// Borderline pairs (|r^2 - rc^2| within the single-precision uncertainty) are decided exactly as the
// CPU oracle decides them: in double, from the untouched input coordinates, lower index first.
static __device__ __noinline__ bool exactInRange(const double* posd, int origI, int origJ, const ExactParams* p) {
    double d[3];
    deltaExact(posd+3*(size_t) min(origI, origJ), posd+3*(size_t) max(origI, origJ), p->boxd, p->periodic != 0, d);
    return !(dot3Exact(d) > p->cutoff2d);
}

template <typename R, int NSUB> struct SubsetTable {
    // lambda and energy tables indexed by the j atom's subset (the i subset is fixed per lane)
    R lamC[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS], lamV[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS];
    R eC[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS], eV[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS];
    double EC[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS], EV[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS];
};

// ---- the staging area: 32 slots of (position+charge, parameters, j index + exclusion mask), addressed
// through explicit 32-bit shared-memory addresses that live in registers.  (Left to itself the compiler
// re-derives slot addresses from %tid and the shared window base in every step of the pair loop.)
__device__ __forceinline__ unsigned opaque(unsigned x) {
    asm volatile("mov.u32 %0, %0;" : "+r"(x));
    return x;
}
template <typename R> struct Stage;
template <> struct Stage<float> {
    enum { STRIDE = 16, PRM = 512, AUX = 1024, BYTES = 1536 };
    static __device__ __forceinline__ float4 ld4(unsigned ad) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ad) : "memory");
        return v;
    }
    static __device__ __forceinline__ void st4(unsigned ad, const float4& v) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(ad), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
};
template <> struct Stage<double> {
    enum { STRIDE = 32, PRM = 1024, AUX = 2048, BYTES = 3072 };
    static __device__ __forceinline__ double4 ld4(unsigned ad) {
        double4 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ad) : "memory");
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(ad+16) : "memory");
        return v;
    }
    static __device__ __forceinline__ void st4(unsigned ad, const double4& v) {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(ad), "d"(v.x), "d"(v.y) : "memory");
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(ad+16), "d"(v.z), "d"(v.w) : "memory");
    }
};
__device__ __forceinline__ unsigned ldSharedU32(unsigned ad) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ad) : "memory");
    return v;
}
__device__ __forceinline__ void stSharedV2U32(unsigned ad, unsigned x, unsigned y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(ad), "r"(x), "r"(y) : "memory");
}

// ---- one pair: Coulomb and van der Waals terms as (energy, dE/dr * r); the common 1/r^2 of the force is
// returned separately and applied once, to the lambda-weighted sum.
template <typename R, int KIND>
__device__ __forceinline__ void pairMath(const DirectConsts<R>& c, R r2, R qq, R sigI, R epsI, R c6I, R sigJ, R epsJ, R c6J,
                                         R& fc, R& fv, R& ec, R& ev, R& invR2) {
    const R invR = Real<R>::rsqrt(r2);
    invR2 = invR*invR;
    const R r = r2*invR;
    if (KIND == KIND_EWALD || KIND == KIND_LJPME) {
        R ex, erfcAR;
        Real<R>::erfcTerms(c, r, r2, ex, erfcAR);
        const R qqr = qq*invR;
        ec = qqr*erfcAR;
        fc = qqr*fma(c.twoAlphaOverSqrtPi*r, ex, erfcAR);
    }
    else if (KIND == KIND_RF) {
        ec = qq*(invR+c.krf*r2-c.crf);
        fc = qq*(invR-c.twoKrf*r2);
    }
    else {
        ec = qq*invR;
        fc = ec;
    }
    const R sig = sigI+sigJ;
    const R eps = epsI*epsJ;
    R s2 = sig*invR;
    s2 *= s2;
    const R s6 = s2*s2*s2;
    fv = eps*(12*s6-6)*s6;
    ev = eps*(s6-1)*s6;
    if (KIND == KIND_LJPME) {
        const R c6 = c6I*c6J;
        const R dar2 = c.dalpha2*r2, dar4 = dar2*dar2;
        const R exd = Real<R>::expNeg(dar2);
        const R invR6 = invR2*invR2*invR2;
        const R pe = 1+dar2+(R) 0.5*dar4;
        ev += c6*invR6*(1-exd*pe);
        fv += 6*c6*invR6*(1-exd*(pe+dar4*dar2*((R) 1/(R) 6)));
        R sc2 = sig*sig;
        const R sc6 = sc2*sc2*sc2*c.invCut6;
        ev += eps*(1-sc6)*sc6-c6*c.invCut6*c.dispShiftExp;
    }
    if (KIND != KIND_PLAIN && KIND != KIND_LJPME) {
        if (r > c.switchDist) {             // +inf when the switching function is off
            const R x = (r-c.switchDist)*c.invSwitchWidth;
            const R sw = 1+x*x*x*(-10+x*(15-x*6));
            const R dsw = x*x*(-30+x*(60-x*30))*c.invSwitchWidth;
            fv = sw*fv-ev*dsw*r;
            ev *= sw;
        }
    }
}

// what the rare out-of-line paths need, by value (taking the address of the kernel's parameter block
// would copy all of it to the stack)
struct DirectOut {
    const double* posd;
    const int* sortedAtom;
    const StepParams* sp;
    long long* forceSorted;
    int paddedAtoms;
    double* energyTable;            // this CTA's replica of the [slice][term] table
    int2* pairDump;
    unsigned long long* pairDumpCount;
    long long pairDumpCapacity;
};

// The pairs of one chunk that fell inside the band around the cutoff (about 1e-5 of all pairs): decided
// exactly, then evaluated here and added with atomics, so that the pair loop itself carries no call.
template <typename R, int KIND, int PBC, bool ENERGY, bool DUMP>
static __device__ __noinline__ void bandPairs(DirectOut o, DirectConsts<R> c, unsigned bandMask, int lane, int ki, int origI,
                                              typename Real<R>::vec4 pi, typename Real<R>::vec4 si4, int perPair, const unsigned char* stage) {
    typedef typename Real<R>::vec4 vec4;
    const int copy = lane>>3, k = lane&7;
    const BoxD boxd = o.sp->exact.boxd;
    while (bandMask) {
        const int s = __ffs(bandMask)-1;
        bandMask &= bandMask-1;
        const int slot = (copy<<3)|((k+s)&7);
        const vec4 q = *(const vec4*) (stage+slot*Stage<R>::STRIDE);
        const vec4 prm = *(const vec4*) (stage+Stage<R>::PRM+slot*Stage<R>::STRIDE);
        const int j = *(const int*) (stage+Stage<R>::AUX+slot*Stage<R>::STRIDE);
        const int origJ = o.sortedAtom[j];
        if (!exactInRange(o.posd, origI, origJ, &o.sp->exact))
            continue;
        R dx = pi.x-q.x, dy = pi.y-q.y, dz = pi.z-q.z;
        if (PBC == PBC_SHIFT && perPair) {
            dx -= (R) boxd.ax*Real<R>::rint(dx/(R) boxd.ax);
            dy -= (R) boxd.by*Real<R>::rint(dy/(R) boxd.by);
            dz -= (R) boxd.cz*Real<R>::rint(dz/(R) boxd.cz);
        }
        else if (PBC == PBC_PAIR_TRICLINIC) {
            R m = Real<R>::rint(dz/(R) boxd.cz);
            dx -= m*(R) boxd.cx; dy -= m*(R) boxd.cy; dz -= m*(R) boxd.cz;
            m = Real<R>::rint(dy/(R) boxd.by);
            dx -= m*(R) boxd.bx; dy -= m*(R) boxd.by;
            m = Real<R>::rint(dx/(R) boxd.ax);
            dx -= m*(R) boxd.ax;
        }
        const R r2 = dx*dx+dy*dy+dz*dz;
        R fc, fv, ec, ev, invR2;
        pairMath<R, KIND>(c, r2, pi.w*q.w, si4.x, si4.y, si4.w, prm.x, prm.y, prm.w, fc, fv, ec, ev, invR2);
        const int slice = sliceOf((int) si4.z, (int) prm.z);
        const R f = ((R) o.sp->lambdaV[slice]*fv+(R) o.sp->lambdaC[slice]*fc)*invR2;
        const long long fx = Real<R>::fixed(f*dx), fy = Real<R>::fixed(f*dy), fz = Real<R>::fixed(f*dz);
        addFixed(&o.forceSorted[ki], fx); addFixed(&o.forceSorted[o.paddedAtoms+ki], fy); addFixed(&o.forceSorted[2*o.paddedAtoms+ki], fz);
        addFixed(&o.forceSorted[j], -fx); addFixed(&o.forceSorted[o.paddedAtoms+j], -fy); addFixed(&o.forceSorted[2*o.paddedAtoms+j], -fz);
        if (ENERGY) {
            atomicAdd(&o.energyTable[2*slice], (double) ec);
            atomicAdd(&o.energyTable[2*slice+1], (double) ev);
        }
        if (DUMP) {
            unsigned long long idx = atomicAdd(o.pairDumpCount, 1ull);
            if ((long long) idx < o.pairDumpCapacity)
                o.pairDump[idx] = make_int2(min(origI, origJ), max(origI, origJ));
        }
    }
}

template <typename R, int KIND, int PBC, int NSUB, bool ENERGY, bool DUMP>
__global__ void __launch_bounds__(32, sizeof(R) == 4 ? (NSUB > 2 || NSUB == 0 ? DIRECT_CTAS_PER_SM/2 : DIRECT_CTAS_PER_SM) : 8) k_direct(DirectArgs a) {
    constexpr int NS = NSUB > 0 ? NSUB : SNB_MAX_SUBSETS;
    constexpr bool TWO = NSUB == 2;         // subsets 0/1 travel as the floats 0.0/1.0: lambdas and energies by FMA
    typedef typename Real<R>::vec4 vec4;
    typedef Stage<R> St;
    // one warp per CTA: the staging area is addressed by the slot alone
    __shared__ __align__(256) unsigned char sStage[St::BYTES];
    const int lane = threadIdx.x;
    const int copy = lane>>3, k = lane&7;
    const unsigned full = 0xffffffffu;
    const DirectParams& p = a.p;
    const DirectConsts<R>& c = Real<R>::consts(a);
    const int numSub = NSUB > 0 ? NSUB : a.numSubsets;
    const int numItems = a.counters[1];
    const BoxD boxd = a.sp->exact.boxd;
    const R bAx = (R) boxd.ax, bBy = (R) boxd.by, bCz = (R) boxd.cz;
    const R iAx = (R) (1.0/boxd.ax), iBy = (R) (1.0/boxd.by), iCz = (R) (1.0/boxd.cz);
    const unsigned kBit = opaque(1u<<k);
    const int srcLane = (int) opaque((unsigned) ((copy<<3)|((k+1)&7)));     // the j accumulators move one lane down inside the copy
    const unsigned stageBase = (unsigned) __cvta_generic_to_shared(sStage);
    const unsigned myAddr = opaque(stageBase+lane*St::STRIDE);
    // slot of step s = 8*copy + (k+s)%8: address = copyBase | ((kOff + s*STRIDE) & (7*STRIDE)), two instructions
    const unsigned copyBase = opaque(stageBase+(copy<<3)*St::STRIDE), kOff = opaque(k*St::STRIDE);
    (void) p;
    // slice energies of this thread over all its work items, keyed by (own subset, j subset), kept in
    // shared memory (they are touched once per work item; registers are worth more in the pair loop);
    // reduced over the warp once, at the end of the kernel
    constexpr int NA = NS > 4 ? 1 : NS;
    __shared__ double sAccC[NA*NA][32], sAccV[NA*NA][32];
    if (ENERGY && NSUB > 0) {
#pragma unroll
        for (int s = 0; s < NA*NA; s++)
            sAccC[s][lane] = sAccV[s][lane] = 0.0;
    }
    for (;;) {
        int item = 0;
        if (lane == 0)
            item = atomicAdd(&a.counters[3], 1);
        item = __shfl_sync(full, item, 0);
        if (item >= numItems)
            break;
        const int4 wi = a.workItems[item];
        const int O = wi.x, I = O>>2;
        const bool perPair = PBC == PBC_SHIFT && wi.w != 0;    // octet too wide for one image per j atom
        const int ki = O*8+k;
        const vec4 pi = Real<R>::pos(a)[ki];
        const vec4 si4 = Real<R>::prm(a)[ki];
        const int subI = (int) si4.z;           // subsets travel as exact small reals, -1 = padding lane
        const bool iValid = subI >= 0;
        const int origI = a.sortedAtom[ki];
        const double cIx = a.blockCenter[3*I], cIy = a.blockCenter[3*I+1], cIz = a.blockCenter[3*I+2];
        const float4 oc = a.octetCenter[O];
        // a padding lane can never be "in range": its bounds are negative
        const R myCutHi = iValid ? c.cutoffHi : (R) -1, myCutLo = iValid ? c.cutoffLo : (R) -1;
        SubsetTable<R, NSUB> t;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            int sl = sliceOf(max(subI, 0), s);
            t.lamC[s] = s < numSub ? (R) a.sp->lambdaC[sl] : (R) 0;
            t.lamV[s] = s < numSub ? (R) a.sp->lambdaV[sl] : (R) 0;
            t.EC[s] = t.EV[s] = 0.0;
        }
        const R dLamC = TWO ? t.lamC[1]-t.lamC[0] : (R) 0, dLamV = TWO ? t.lamV[1]-t.lamV[0] : (R) 0;
        double Fx = 0, Fy = 0, Fz = 0;
        for (int ch = 0; ch < wi.z; ch++) {
            // ---- gather and stage the chunk's j atoms -------------------------------------
            const int2 ent = a.jList[(size_t) (wi.y+ch)*32+lane];
            vec4 pj, sj4;
            pj.x = pj.y = pj.z = pj.w = 0; sj4.x = sj4.y = sj4.z = sj4.w = 0;
            if (ent.x >= 0) {
                pj = Real<R>::pos(a)[ent.x];
                sj4 = Real<R>::prm(a)[ent.x];
                const int J = ent.x>>5;
                R ox = (R) (a.blockCenter[3*J]-cIx), oy = (R) (a.blockCenter[3*J+1]-cIy), oz = (R) (a.blockCenter[3*J+2]-cIz);
                pj.x += ox; pj.y += oy; pj.z += oz;
                if (PBC == PBC_SHIFT && !perPair) {
                    // one image per j atom, chosen relative to the octet's centre (valid while
                    // halfBox - cutoff exceeds the octet's half extent; the list build flags the rest)
                    pj.x -= bAx*Real<R>::rint((pj.x-(R) oc.x)*iAx);
                    pj.y -= bBy*Real<R>::rint((pj.y-(R) oc.y)*iBy);
                    pj.z -= bCz*Real<R>::rint((pj.z-(R) oc.z)*iCz);
                }
            }
            __syncwarp();
            St::st4(myAddr, pj);
            St::st4(myAddr+St::PRM, sj4);
            stSharedV2U32(myAddr+St::AUX, (unsigned) ent.x, (unsigned) ent.y);
            __syncwarp();
            R fix = 0, fiy = 0, fiz = 0, fjx = 0, fjy = 0, fjz = 0;
            R eCt = 0, eVt = 0, eC1 = 0, eV1 = 0;       // TWO: totals and the j-subset-1 parts
            unsigned bandMask = 0;
#pragma unroll
            for (int s = 0; s < NS; s++)
                t.eC[s] = t.eV[s] = 0;
            // ---- 8 steps of 32 pairs --------------------------------------------------------
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const unsigned ad = copyBase|((kOff+s*St::STRIDE)&(7*St::STRIDE));
                const vec4 q = St::ld4(ad);
                const unsigned mj = ldSharedU32(ad+St::AUX+4);
                R dx = pi.x-q.x, dy = pi.y-q.y, dz = pi.z-q.z;
                if (PBC == PBC_SHIFT && perPair) {
                    dx -= bAx*Real<R>::rint(dx*iAx);
                    dy -= bBy*Real<R>::rint(dy*iBy);
                    dz -= bCz*Real<R>::rint(dz*iCz);
                }
                else if (PBC == PBC_PAIR_TRICLINIC) {
                    R m = Real<R>::rint(dz*iCz);
                    dx -= m*(R) boxd.cx; dy -= m*(R) boxd.cy; dz -= m*bCz;
                    m = Real<R>::rint(dy*iBy);
                    dx -= m*(R) boxd.bx; dy -= m*bBy;
                    m = Real<R>::rint(dx*iAx);
                    dx -= m*bAx;
                }
                const R r2 = dx*dx+dy*dy+dz*dz;
                // the hot path takes the pairs safely inside the cutoff; those inside the band around it (about
                // 1e-5 of all pairs) are only noted, and settled exactly after the eight steps (bandPairs)
                const bool ok = !(mj&kBit);
                bool in;
                if (KIND != KIND_PLAIN) {
                    in = ok && r2 < myCutLo;
                    if (ok && !(r2 < myCutLo) && r2 < myCutHi)
                        bandMask |= 1u<<s;
                }
                else
                    in = ok && iValid;
                if (in) {
                    const vec4 prm = St::ld4(ad+St::PRM);
                    R fc, fv, ec, ev, invR2;
                    pairMath<R, KIND>(c, r2, pi.w*q.w, si4.x, si4.y, si4.w, prm.x, prm.y, prm.w, fc, fv, ec, ev, invR2);
                    R lc, lv;
                    if (NSUB == 1) {
                        lc = t.lamC[0]; lv = t.lamV[0];
                    }
                    else if (TWO) {
                        lc = fma(prm.z, dLamC, t.lamC[0]);
                        lv = fma(prm.z, dLamV, t.lamV[0]);
                    }
                    else if (NSUB > 0) {
                        lc = t.lamC[0]; lv = t.lamV[0];
#pragma unroll
                        for (int u = 1; u < NS; u++) {
                            lc = prm.z == (R) u ? t.lamC[u] : lc;
                            lv = prm.z == (R) u ? t.lamV[u] : lv;
                        }
                    }
                    else {
                        const int sl = sliceOf(subI, (int) prm.z);
                        lc = (R) a.sp->lambdaC[sl]; lv = (R) a.sp->lambdaV[sl];
                    }
                    const R f = (lv*fv+lc*fc)*invR2;
                    fix += f*dx; fiy += f*dy; fiz += f*dz;
                    fjx -= f*dx; fjy -= f*dy; fjz -= f*dz;
                    if (ENERGY) {
                        if (TWO) {
                            eCt += ec; eVt += ev;
                            eC1 = fma(prm.z, ec, eC1); eV1 = fma(prm.z, ev, eV1);
                        }
                        else if (NSUB > 0) {
#pragma unroll
                            for (int u = 0; u < NS; u++) {
                                t.eC[u] += prm.z == (R) u ? ec : (R) 0;
                                t.eV[u] += prm.z == (R) u ? ev : (R) 0;
                            }
                        }
                        else {
                            // many subsets: straight to the global table (rare configuration)
                            double* table = a.energyBuf+(size_t) (blockIdx.x%ENERGY_REPLICAS)*2*SNB_MAX_SLICES;
                            atomicAdd(&table[2*sliceOf(subI, (int) prm.z)], (double) ec);
                            atomicAdd(&table[2*sliceOf(subI, (int) prm.z)+1], (double) ev);
                        }
                    }
                    if (DUMP) {
                        const int origJ = a.sortedAtom[(int) ldSharedU32(ad+St::AUX)];
                        unsigned long long idx = atomicAdd(a.pairDumpCount, 1ull);
                        if ((long long) idx < a.pairDumpCapacity)
                            a.pairDump[idx] = make_int2(min(origI, origJ), max(origI, origJ));
                    }
                }
                fjx = __shfl_sync(full, fjx, srcLane);
                fjy = __shfl_sync(full, fjy, srcLane);
                fjz = __shfl_sync(full, fjz, srcLane);
            }
            if (KIND != KIND_PLAIN && __any_sync(full, bandMask != 0u)) {
                if (bandMask) {
                    DirectOut o;
                    o.posd = a.posd; o.sortedAtom = a.sortedAtom; o.sp = a.sp; o.forceSorted = a.forceSorted; o.paddedAtoms = a.paddedAtoms;
                    o.energyTable = a.energyBuf+(size_t) (blockIdx.x%ENERGY_REPLICAS)*2*SNB_MAX_SLICES;
                    o.pairDump = a.pairDump; o.pairDumpCount = a.pairDumpCount; o.pairDumpCapacity = a.pairDumpCapacity;
                    bandPairs<R, KIND, PBC, ENERGY, DUMP>(o, c, bandMask, lane, ki, origI, pi, si4, perPair ? 1 : 0, sStage);
                }
            }
            if (ent.x >= 0) {
                addFixed(&a.forceSorted[ent.x], Real<R>::fixed(fjx));
                addFixed(&a.forceSorted[a.paddedAtoms+ent.x], Real<R>::fixed(fjy));
                addFixed(&a.forceSorted[2*a.paddedAtoms+ent.x], Real<R>::fixed(fjz));
            }
            Fx += fix; Fy += fiy; Fz += fiz;
            if (ENERGY && TWO) {
                t.EC[0] += eCt-eC1; t.EC[1] += eC1;
                t.EV[0] += eVt-eV1; t.EV[1] += eV1;
            }
            else if (ENERGY && NSUB > 0) {
#pragma unroll
                for (int u = 0; u < NS; u++) {
                    t.EC[u] += t.eC[u];
                    t.EV[u] += t.eV[u];
                }
            }
        }
        // the four copies of the octet: sum their i forces (and energies) onto copy 0
        for (int o = 8; o < 32; o <<= 1) {
            Fx += __shfl_xor_sync(full, Fx, o);
            Fy += __shfl_xor_sync(full, Fy, o);
            Fz += __shfl_xor_sync(full, Fz, o);
        }
        if (iValid && copy == 0) {
            addFixed(&a.forceSorted[ki], toFixedD(Fx));
            addFixed(&a.forceSorted[a.paddedAtoms+ki], toFixedD(Fy));
            addFixed(&a.forceSorted[2*a.paddedAtoms+ki], toFixedD(Fz));
        }
        if (ENERGY && NSUB > 0 && iValid) {
#pragma unroll
            for (int u = 0; u < NS; u++) {
                sAccC[subI*NA+u][lane] += t.EC[u];
                sAccV[subI*NA+u][lane] += t.EV[u];
            }
        }
    }
    if (ENERGY && NSUB > 0) {
        double* table = a.energyBuf+(size_t) (blockIdx.x%ENERGY_REPLICAS)*2*SNB_MAX_SLICES;
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int u = 0; u < NS; u++) {
                double ecs = sAccC[s*NA+u][lane], evs = sAccV[s*NA+u][lane];
                for (int o = 16; o > 0; o >>= 1) {
                    ecs += __shfl_xor_sync(full, ecs, o);
                    evs += __shfl_xor_sync(full, evs, o);
                }
                if (lane == 0 && s < numSub && u < numSub) {
                    if (ecs != 0.0) atomicAdd(&table[2*sliceOf(s, u)], ecs);
                    if (evs != 0.0) atomicAdd(&table[2*sliceOf(s, u)+1], evs);
                }
            }
    }
}

template <int KIND, int PBC, int NSUB>
static void launch4(const DirectArgs& a, cudaStream_t st, int grid) {
    if (a.doublePrecision) {
        // the double-precision mode is a correctness mode: one instantiation (energies always on)
        if (a.pairDump) k_direct<double, KIND, PBC, NSUB, true, true><<<grid/3, 32, 0, st>>>(a);
        else k_direct<double, KIND, PBC, NSUB, true, false><<<grid/3, 32, 0, st>>>(a);
        return;
    }
    if (a.pairDump)
        k_direct<float, KIND, PBC, NSUB, true, true><<<grid, 32, 0, st>>>(a);
    else if (a.needEnergy)
        k_direct<float, KIND, PBC, NSUB, true, false><<<grid, 32, 0, st>>>(a);
    else
        k_direct<float, KIND, PBC, NSUB, false, false><<<grid, 32, 0, st>>>(a);
}

template <int KIND, int PBC>
static void launch3(const DirectArgs& a, cudaStream_t st, int grid) {
    switch (a.numSubsets) {
        case 1: launch4<KIND, PBC, 1>(a, st, grid); break;
        case 2: launch4<KIND, PBC, 2>(a, st, grid); break;
        case 3:
        case 4: launch4<KIND, PBC, 4>(a, st, grid); break;
        default: launch4<KIND, PBC, 0>(a, st, grid); break;
    }
}

template <int KIND>
void launchDirectKind(const DirectArgs& a, cudaStream_t st, int grid) {
    switch (a.p.pbcMode) {
        case PBC_NONE:
            if (KIND == KIND_PLAIN || KIND == KIND_RF) launch3<KIND, PBC_NONE>(a, st, grid);
            break;
        case PBC_SHIFT:
        case PBC_PAIR_ORTHO:
            // orthorhombic box: one image per j atom, or per-pair wrapping for the (few) octets whose
            // bounding box is too wide for that -- decided per work item
            if (KIND != KIND_PLAIN) launch3<KIND, PBC_SHIFT>(a, st, grid);
            break;
        default:
            if (KIND != KIND_PLAIN) launch3<KIND, PBC_PAIR_TRICLINIC>(a, st, grid);
            break;
    }
}

#ifdef SNB_DIRECT_KIND
// one translation unit per interaction kind, compiled in parallel (see build.py)
template void launchDirectKind<SNB_DIRECT_KIND>(const DirectArgs&, cudaStream_t, int);
#endif
