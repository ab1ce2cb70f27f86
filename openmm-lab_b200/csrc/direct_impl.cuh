// Direct-space tile kernel (north-star subsystem 2).
//
// The tile list is a flat array of chunks: 32 compacted j atoms against one octet of i atoms.  Every warp
// (one-warp CTAs) takes an equal, contiguous range of chunks -- chunks cost the same, so the static split
// balances -- and streams through it:
//   * the octet is held four times over in the warp (lane = 8*copy + k, atom k of the octet); copy c owns
//     j slots 8c..8c+7 of the chunk, so a chunk is 8 steps of 32 pairs.  In step s lane (c,k) works on slot
//     8c + (k^s): every step touches 32 different pairs and the slot's address is the lane's own address
//     XOR a constant; the j-force accumulators follow their slot between the 8 lanes of a copy by butterfly
//     shuffles (lane masks s^(s+1), then 7) and are home after 8 steps, so no reduction tree is needed; the
//     four copies' i forces are summed by shuffle when the octet changes;
//   * a chunk's j atoms (position+charge, parameters, block centre) are gathered two chunks ahead: the list
//     entry by a plain load, the atom data by cp.async straight into the other half of a double-buffered
//     shared-memory stage, so the gather's two dependent global-memory latencies hide behind the pair
//     arithmetic of the current chunk; one fix-up pass shifts the staged atoms into the octet's frame
//     (exact float block-centre differences, one periodic image per j atom where the box allows it);
//   * the pair arithmetic is branch-free in single precision: out-of-range / excluded lanes compute on a
//     harmless r^2 and are masked out by selects (nearly every step has some lane in range, so a branch
//     saves no issue slots and only blocks the scheduler from overlapping the steps).
// Pair arithmetic is FP32 (MUFU rsqrt + Newton step, MUFU ex2, polynomial erfc); every chunk's force
// sums are converted to 2^32 fixed point and all further accumulation is in 64-bit integers (order and
// partition independent => bit-reproducible, also across GPUs); slice energies are accumulated per j subset in FP32 per chunk and FP64 beyond.
//
// Pair arithmetic follows the reference's Reference platform:
//   Ewald/PME/LJPME direct:  platforms/reference/src/ReferenceSlicedLJCoulombIxn.cpp:351-429
//   cutoff / no cutoff:      ReferenceSlicedLJCoulombIxn.cpp:553-611 (reaction field :64-65)
// (behavioural cross-reference on the GPU side: platforms/common/src/kernels/coulombLennardJones.cc:1-124).
#include "kernels.cuh"

#include <type_traits>

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
    // Ewald direct-space Coulomb term (one MUFU.EX2, one MUFU.RCP, twelve FMA-pipe instructions): energy
    // qq erfc(ar)/r and (dE/dr) r = qq/r (erfc(ar) + 2 a r/sqrt(pi) exp(-a^2 r^2)), with the common factor
    // qq/r exp(-a^2 r^2) taken out of both (erfc = exp(-x^2) erfcx)
    static __device__ __forceinline__ void ewaldCoulomb(const DirectConsts<float>& c, float r, float r2, float qqr, float& ec, float& fc) {
        const float g = qqr*fastExp2(r2*c.negAlpha2Log2e);
        const float p = erfcxPolyT(fastRcp(fmaf(r, c.halfAlpha, 1.0f)));
        ec = g*p;
        fc = g*fmaf(c.twoAlphaOverSqrtPi, r, p);
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
    static __device__ __forceinline__ void ewaldCoulomb(const DirectConsts<double>& c, double r, double r2, double qqr, double& ec, double& fc) {
        const double erfcAR = erfc(c.alpha*r);
        ec = qqr*erfcAR;
        fc = qqr*fma(c.twoAlphaOverSqrtPi*r, exp(-c.alpha2*r2), erfcAR);
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
// One stage buffer: pos[32] | prm[32] | ctr[32] (float4 block centres) | aux[32] (j, mask, -, -); slot
// stride = sizeof(vec4) for pos/prm, 16 for ctr and aux (in single precision every section of a slot then
// sits at a constant offset from the slot's position: one address per step).
template <typename R> struct Stage;
template <> struct Stage<float> {
    enum { STRIDE = 16, PRM = 512, CTR = 1024, AUX = 1536, AUXS = 16, BYTES = 2048 };
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
    enum { STRIDE = 32, PRM = 1024, CTR = 2048, AUX = 2560, AUXS = 16, BYTES = 3072 };
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
__device__ __forceinline__ float4 ldSharedF4(unsigned ad) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ad) : "memory");
    return v;
}
// 16-byte asynchronous copy global -> shared (LDGSTS): no registers, completion tracked per thread in groups
__device__ __forceinline__ void cpAsync16(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsync8(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsync4(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
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
template <typename R, int KIND, bool SWITCH>
__device__ __forceinline__ void pairMath(const DirectConsts<R>& c, R r2, R qq, R sigI, R epsI, R c6I, R sigJ, R epsJ, R c6J,
                                         R& fc, R& fv, R& ec, R& ev, R& invR2) {
    const R invR = Real<R>::rsqrt(r2);
    invR2 = invR*invR;
    const R r = r2*invR;
    if (KIND == KIND_EWALD || KIND == KIND_LJPME)
        Real<R>::ewaldCoulomb(c, r, r2, qq*invR, ec, fc);
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
    const R es6 = eps*s6;
    fv = es6*fma((R) 12, s6, (R) -6);      // eps (12 s^12 - 6 s^6)
    ev = fma(es6, s6, -es6);               // eps (s^12 - s^6)
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
    if (SWITCH) {
        // branch-free: below the switching distance x = 0, hence sw = 1 and dsw = 0
        const R x = max((R) 0, (r-c.switchDist)*c.invSwitchWidth);
        const R sw = 1+x*x*x*(-10+x*(15-x*6));
        const R dsw = x*x*(-30+x*(60-x*30))*c.invSwitchWidth;
        fv = sw*fv-ev*dsw*r;
        ev *= sw;
    }
}

// Box lengths (and, for triclinic boxes, the off-diagonal components) the wrapping needs, in registers.  The image
// shift of the fix-up pass uses every length as hi + lo: hi on the block centres' 2^-10 nm lattice, lo the remainder.
template <typename R> struct WrapBox {
    R ax, by, cz, iax, iby, icz, bx, cx, cy;
    R axHi, byHi, czHi, axLo, byLo, czLo;
};
template <typename R> __device__ __forceinline__ WrapBox<R> makeWrapBox(const BoxD& b) {
    WrapBox<R> w;
    w.ax = (R) b.ax; w.by = (R) b.by; w.cz = (R) b.cz;
    w.iax = (R) (1.0/b.ax); w.iby = (R) (1.0/b.by); w.icz = (R) (1.0/b.cz);
    w.bx = (R) b.bx; w.cx = (R) b.cx; w.cy = (R) b.cy;
    const double hx = rint(b.ax*1024.0)*(1.0/1024.0), hy = rint(b.by*1024.0)*(1.0/1024.0), hz = rint(b.cz*1024.0)*(1.0/1024.0);
    w.axHi = (R) hx; w.byHi = (R) hy; w.czHi = (R) hz;
    w.axLo = (R) (b.ax-hx); w.byLo = (R) (b.by-hy); w.czLo = (R) (b.cz-hz);
    return w;
}
// Displacement j -> i of one pair and its squared length.  Every rounding is pinned (explicit intrinsics are never
// re-contracted), because the pair loop and the band path both classify a pair from this r^2 and must agree
// bit for bit.  WRAP: 0 none (the staged j atom already is the right image), 1 orthorhombic per pair, 2 triclinic.
template <int WRAP> __device__ __forceinline__ float pairDelta(const float4& pi, const float4& q, const WrapBox<float>& w, float& dx, float& dy, float& dz) {
    dx = __fsub_rn(pi.x, q.x); dy = __fsub_rn(pi.y, q.y); dz = __fsub_rn(pi.z, q.z);
    if (WRAP == 1) {
        dx = __fmaf_rn(-w.ax, rintf(__fmul_rn(dx, w.iax)), dx);
        dy = __fmaf_rn(-w.by, rintf(__fmul_rn(dy, w.iby)), dy);
        dz = __fmaf_rn(-w.cz, rintf(__fmul_rn(dz, w.icz)), dz);
    }
    else if (WRAP == 2) {
        float m = rintf(__fmul_rn(dz, w.icz));
        dx = __fmaf_rn(-m, w.cx, dx); dy = __fmaf_rn(-m, w.cy, dy); dz = __fmaf_rn(-m, w.cz, dz);
        m = rintf(__fmul_rn(dy, w.iby));
        dx = __fmaf_rn(-m, w.bx, dx); dy = __fmaf_rn(-m, w.by, dy);
        m = rintf(__fmul_rn(dx, w.iax));
        dx = __fmaf_rn(-m, w.ax, dx);
    }
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
template <int WRAP> __device__ __forceinline__ double pairDelta(const double4& pi, const double4& q, const WrapBox<double>& w, double& dx, double& dy, double& dz) {
    dx = __dsub_rn(pi.x, q.x); dy = __dsub_rn(pi.y, q.y); dz = __dsub_rn(pi.z, q.z);
    if (WRAP == 1) {
        dx = __fma_rn(-w.ax, rint(__dmul_rn(dx, w.iax)), dx);
        dy = __fma_rn(-w.by, rint(__dmul_rn(dy, w.iby)), dy);
        dz = __fma_rn(-w.cz, rint(__dmul_rn(dz, w.icz)), dz);
    }
    else if (WRAP == 2) {
        double m = rint(__dmul_rn(dz, w.icz));
        dx = __fma_rn(-m, w.cx, dx); dy = __fma_rn(-m, w.cy, dy); dz = __fma_rn(-m, w.cz, dz);
        m = rint(__dmul_rn(dy, w.iby));
        dx = __fma_rn(-m, w.bx, dx); dy = __fma_rn(-m, w.by, dy);
        m = rint(__dmul_rn(dx, w.iax));
        dx = __fma_rn(-m, w.ax, dx);
    }
    return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

// Fix-up of one staged j atom: from its own block's frame into the i block's, and -- `image` -- onto the periodic
// image nearest the octet's centre (valid while halfBox - cutoff exceeds the octet's half extent; the list build
// flags the other octets for per-pair wrapping).  Block centres sit on a 2^-10 nm lattice, and so does the hi part
// of every box length: centre difference - m*hi is EXACT in float (an integer below 2^24 over 1024), a few nm at most
// for any pair near the cutoff, so adding it to the local coordinate rounds at that magnitude only; m*lo (below
// 2^-11 nm per image) follows by FMA.  r^2 thus keeps an absolute error of ~5e-7 nm^2 whatever the box size, which
// is what the narrow cutoff band relies on.
template <typename R> __device__ __forceinline__ void fixupShift(typename Real<R>::vec4& pj, const float4& cJ, const float4& cI, const float4& oc,
                                                                 bool image, const WrapBox<R>& wb) {
    const R sx = (R) (cJ.x-cI.x), sy = (R) (cJ.y-cI.y), sz = (R) (cJ.z-cI.z);
    if (!image) {
        pj.x += sx; pj.y += sy; pj.z += sz;
        return;
    }
    const R mx = Real<R>::rint((pj.x+sx-(R) oc.x)*wb.iax), my = Real<R>::rint((pj.y+sy-(R) oc.y)*wb.iby),
            mz = Real<R>::rint((pj.z+sz-(R) oc.z)*wb.icz);
    pj.x = fma(-mx, wb.axLo, pj.x+fma(-mx, wb.axHi, sx));
    pj.y = fma(-my, wb.byLo, pj.y+fma(-my, wb.byHi, sy));
    pj.z = fma(-mz, wb.czLo, pj.z+fma(-mz, wb.czHi, sz));
}

// Chunks of the tile list this evaluation works on (0 when the list overflowed its buffers: the host discards the pass and
// repeats it).  List reuse: an evaluation that kept the list finds its length in listState[0] (device-resident between
// evaluations; the counters are zeroed every evaluation), one that rebuilt it leaves the new length there.
__device__ __forceinline__ int listChunks(const DirectArgs& a) {
    int n = a.counters[0];
    if (a.listState) {
        if (a.counters[SNB_COUNTER_REBUILD] == 0)
            n = a.listState[0];
        if (blockIdx.x == 0 && threadIdx.x == 0 && !a.counters[2]) {
            a.listState[0] = n;
            a.counters[0] = n;          // (the host reads the length from the counters)
        }
    }
    return a.counters[2] ? 0 : min(n, a.maxChunks);
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
template <typename R, int KIND, int PBC, bool ENERGY, bool DUMP, bool SWITCH>
static __device__ __noinline__ void bandPairs(DirectOut o, DirectConsts<R> c, unsigned bandMask, int lane, int ki, int origI,
                                              typename Real<R>::vec4 pi, typename Real<R>::vec4 si4, int perPair, const unsigned char* stage) {
    typedef typename Real<R>::vec4 vec4;
    const int copy = lane>>3, k = lane&7;
    const BoxD boxd = o.sp->exact.boxd;
    while (bandMask) {
        const int s = __ffs(bandMask)-1;
        bandMask &= bandMask-1;
        const int slot = (copy<<3)|(k^s);
        const vec4 q = *(const vec4*) (stage+slot*Stage<R>::STRIDE);
        const vec4 prm = *(const vec4*) (stage+Stage<R>::PRM+slot*Stage<R>::STRIDE);
        const int j = *(const int*) (stage+Stage<R>::AUX+slot*Stage<R>::AUXS);
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
        pairMath<R, KIND, SWITCH>(c, r2, pi.w*q.w, si4.x, si4.y, si4.w, prm.x, prm.y, prm.w, fc, fv, ec, ev, invR2);
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

// (packed kernel, direct_packed.cu)  The pairs of one chunk that fell inside the band around the cutoff: the pair loop only
// notes that there is one; here the lane finds them again (same r^2, bit for bit: pairDelta), decides each exactly,
// evaluates it and adds it with atomics, so that the pair loop itself carries neither a call nor a per-step flag.
template <typename R, int KIND, int PBC, bool ENERGY, bool DUMP, bool SWITCH>
static __device__ __noinline__ void bandRescan(DirectOut o, DirectConsts<R> c, int lane, int ki, int origI, R cutLo, R cutHi,
                                              typename Real<R>::vec4 pi, typename Real<R>::vec4 si4, int perPair, const unsigned char* stage) {
    typedef typename Real<R>::vec4 vec4;
    const int copy = lane>>3, k = lane&7;
    const WrapBox<R> wb = makeWrapBox<R>(o.sp->exact.boxd);
    for (int s = 0; s < 8; s++) {
        const int slot = (copy<<3)|(k^s);
        const vec4 q = *(const vec4*) (stage+slot*Stage<R>::STRIDE);
        const int2 aux = *(const int2*) (stage+Stage<R>::AUX+slot*Stage<R>::AUXS);
        R dx, dy, dz, r2;
        if (PBC == PBC_PAIR_TRICLINIC) r2 = pairDelta<2>(pi, q, wb, dx, dy, dz);
        else if (PBC == PBC_SHIFT && perPair) r2 = pairDelta<1>(pi, q, wb, dx, dy, dz);
        else r2 = pairDelta<0>(pi, q, wb, dx, dy, dz);
        if (((unsigned) aux.y>>k)&1u || r2 < cutLo || !(r2 < cutHi))
            continue;
        const vec4 prm = *(const vec4*) (stage+Stage<R>::PRM+slot*Stage<R>::STRIDE);
        const int j = aux.x;
        const int origJ = o.sortedAtom[j];
        if (!exactInRange(o.posd, origI, origJ, &o.sp->exact))
            continue;
        R fc, fv, ec, ev, invR2;
        pairMath<R, KIND, SWITCH>(c, r2, pi.w*q.w, si4.x, si4.y, si4.w, prm.x, prm.y, prm.w, fc, fv, ec, ev, invR2);
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

template <typename R, int KIND, int PBC, int NSUB, bool ENERGY, bool DUMP, bool SWITCH>
__global__ void __launch_bounds__(32, sizeof(R) == 4 ? (NSUB > 2 || NSUB == 0 ? DIRECT_WIDE_CTAS_PER_SM : DIRECT_CTAS_PER_SM) : 8) k_direct(DirectArgs a) {
    constexpr int NS = NSUB > 0 ? NSUB : SNB_MAX_SUBSETS;
    constexpr bool TWO = NSUB == 2;         // subsets 0/1 travel as the floats 0.0/1.0: lambdas and energies by FMA
    constexpr int STEP_UNROLL = DIRECT_UNROLL;
    constexpr bool BRANCHY = sizeof(R) == 8 || NSUB == 0;   // double precision: skip the arithmetic of masked lanes
    // Ewald/PME without switching: a masked lane computes on r^2 = 1e10 nm^2, where exp(-a^2 r^2) = 0 kills the
    // Coulomb term exactly and the Lennard-Jones force is ~1e-43 of a real one -- nothing to select away
    constexpr bool SELF_MASKING = !BRANCHY && KIND == KIND_EWALD && !SWITCH;
    constexpr bool CLEANABLE = !BRANCHY && NSUB > 0 && KIND != KIND_PLAIN;      // the lean copy of the pair loop exists (see `clean`)
    typedef typename Real<R>::vec4 vec4;
    typedef Stage<R> St;
    // one warp per CTA: the two stage buffers are addressed by the slot alone
    __shared__ __align__(256) unsigned char sStage[2][St::BYTES+(256-St::BYTES%256)%256];
    constexpr unsigned BUF = sizeof(sStage[0]);
    const int lane = threadIdx.x;
    const int copy = lane>>3, k = lane&7;
    const unsigned full = 0xffffffffu;
    const DirectConsts<R>& c = Real<R>::consts(a);
    const int numSub = NSUB > 0 ? NSUB : a.numSubsets;
    // this warp's contiguous share of the chunks
    // (values loaded from a uniform address are broadcast through a shuffle so that the compiler KNOWS they are
    // warp-uniform: the loops and branches they steer then carry no divergence bookkeeping)
    // (a list that overflowed its buffers holds unwritten chunks: the host discards this pass and repeats it)
    const int numChunks = __shfl_sync(full, listChunks(a), 0);
    const int per = (numChunks+(int) gridDim.x-1)/(int) gridDim.x;
    const int cBegin = min(numChunks, (int) blockIdx.x*per), cEnd = min(numChunks, cBegin+per);
    const BoxD boxd = a.sp->exact.boxd;
    const R bAx = (R) boxd.ax, bBy = (R) boxd.by, bCz = (R) boxd.cz;
    const R iAx = (R) (1.0/boxd.ax), iBy = (R) (1.0/boxd.by), iCz = (R) (1.0/boxd.cz);
    const WrapBox<R> wb = makeWrapBox<R>(boxd);
    // cutoff band (snb_core.cuh DirectParams): narrow where the j atoms reach the octet's frame through the exact image
    // shift of the fix-up pass, wide for per-pair wrapping and when some block of this evaluation is large
    const bool narrowOk = PBC != PBC_PAIR_TRICLINIC && !(__int_as_float(a.counters[SNB_COUNTER_MAX_EXTENT]) > SNB_NARROW_MAX_EXTENT);
    const R bandHi = narrowOk ? c.cutoffHi : c.cutoffHiWide, bandLo = narrowOk ? c.cutoffLo : c.cutoffLoWide;
    const unsigned kBit = opaque(1u<<k);
    const unsigned stageBase = opaque((unsigned) __cvta_generic_to_shared(&sStage[0][0]));
    // slot of step s = 8*copy + (k^s): address = (buffer + own slot) ^ (s*STRIDE), one instruction (the buffers
    // are aligned to 8 slots, so the XOR never carries out of the copy's eight slots)
    const unsigned ownOff = opaque((unsigned) lane*St::STRIDE);
    // slice energies of this thread over all its chunks, keyed by (own subset, j subset), kept in shared
    // memory (touched once per octet; registers are worth more in the pair loop); reduced over the warp once
    constexpr int NA = NS > 4 ? 1 : NS;
    __shared__ double sAccC[NA*NA][32], sAccV[NA*NA][32];
    if (ENERGY && NSUB > 0) {
#pragma unroll
        for (int s = 0; s < NA*NA; s++)
            sAccC[s][lane] = sAccV[s][lane] = 0.0;
    }

    // Everything the steady-state loop reads from global memory arrives through cp.async: a plain prefetch load holds a
    // register scoreboard for a DRAM round trip, and with six scoreboards per warp some unrelated instruction of the
    // pair loop ends up waiting on it (ncu on the packed kernel: 26% of the warp time in long-scoreboard stalls).
    //   fetchEntries(c): list entry of every lane + the octet of chunk c  -> ring slot c&3           (two chunks ahead)
    //   gather(c):       the 32 j atoms of chunk c (entries already in the ring) -> stage buffer c&1  (one chunk ahead)
    __shared__ __align__(16) int2 sEnt[4][32];
    __shared__ int sOct[4];
    const unsigned entBase = opaque((unsigned) __cvta_generic_to_shared(&sEnt[0][0])+(unsigned) lane*8u);
    const unsigned octBase = opaque((unsigned) __cvta_generic_to_shared(&sOct[0]));
    auto fetchEntries = [&](int chunk) {
        cpAsync8(entBase+(unsigned) (chunk&3)*256u, a.jList+(size_t) chunk*32+lane);
        if (lane == 0)
            cpAsync4(octBase+(unsigned) (chunk&3)*4u, a.chunkOctet+chunk);
    };
    auto gather = [&](int chunk) {
        const unsigned ent = entBase+(unsigned) (chunk&3)*256u;
        const int jRaw = (int) ldSharedU32(ent);
        const int j = max(jRaw, 0);         // padding entries copy atom 0: never in range (mask 0xff)
        const unsigned dst = stageBase+(unsigned) (chunk&1)*BUF;
        const char* ps = (const char*) (Real<R>::pos(a)+j);
        const char* qs = (const char*) (Real<R>::prm(a)+j);
#pragma unroll
        for (int b = 0; b < (int) sizeof(vec4); b += 16) {
            cpAsync16(dst+lane*St::STRIDE+b, ps+b);
            cpAsync16(dst+St::PRM+lane*St::STRIDE+b, qs+b);
        }
        cpAsync16(dst+St::CTR+lane*16, a.blockCenter+(j>>5));
        stSharedV2U32(dst+St::AUX+lane*St::AUXS, (unsigned) jRaw, ldSharedU32(ent+4));
    };

    // ---- the octet in flight
    int curO = -1, ki = 0, origI = -1, subI = -1;
    bool iValid = false, perPair = false;
    vec4 pi, si4;
    pi.x = pi.y = pi.z = pi.w = 0; si4 = pi;
    float4 cI = make_float4(0.f, 0.f, 0.f, 0.f), oc = cI;
    R myCutHi = -1, myCutLo = -1, dLamC = 0, dLamV = 0;
    SubsetTable<R, NSUB> t;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        t.lamC[s] = t.lamV[s] = 0;
        t.EC[s] = t.EV[s] = 0.0;
    }
    // i forces: every chunk's FP32 partial sum is converted to 2^32 fixed point before it is added, so the
    // total does not depend on how the chunks are grouped into warps, GPUs or octet runs (integer sums)
    long long Fx = 0, Fy = 0, Fz = 0;
    auto flushOctet = [&]() {
        // the four copies of the octet: sum their i forces onto copy 0
        for (int o = 8; o < 32; o <<= 1) {
            Fx += __shfl_xor_sync(full, Fx, o);
            Fy += __shfl_xor_sync(full, Fy, o);
            Fz += __shfl_xor_sync(full, Fz, o);
        }
        if (iValid && copy == 0) {
            addFixed(&a.forceSorted[ki], Fx);
            addFixed(&a.forceSorted[a.paddedAtoms+ki], Fy);
            addFixed(&a.forceSorted[2*a.paddedAtoms+ki], Fz);
        }
        if (ENERGY && NSUB > 0 && iValid) {
#pragma unroll
            for (int u = 0; u < NS; u++) {
                sAccC[subI*NA+u][lane] += t.EC[u];
                sAccV[subI*NA+u][lane] += t.EV[u];
            }
        }
        Fx = Fy = Fz = 0;
#pragma unroll
        for (int s = 0; s < NS; s++)
            t.EC[s] = t.EV[s] = 0.0;
    };

    // ---- prologue: entries of the first two chunks, then the atoms of the first
    if (cBegin < cEnd) {
        fetchEntries(cBegin);
        if (cBegin+1 < cEnd)
            fetchEntries(cBegin+1);
        cpAsyncCommit();
        cpAsyncWait<0>();
        __syncwarp();
        gather(cBegin);
        cpAsyncCommit();
    }
    for (int ch = cBegin; ch < cEnd; ch++) {
        const unsigned buf = (unsigned) ch&1u;
        const unsigned sb = stageBase+buf*BUF;
        // the atoms of chunk ch and the entries of chunk ch+1 were requested a whole chunk ago
        cpAsyncWait<0>();
        __syncwarp();                       // ... by every lane; and every lane is done with chunk ch-1
        if (ch+1 < cEnd) {
            gather(ch+1);
            if (ch+2 < cEnd)
                fetchEntries(ch+2);
            cpAsyncCommit();
        }
        const int octCur = __shfl_sync(full, (int) ldSharedU32(octBase+(unsigned) (ch&3)*4u), 0);
        // ---- octet switch
        const int O = octCur&0x3fffffff;
        if (O != curO) {
            if (curO >= 0)
                flushOctet();
            curO = O;
            ki = O*8+k;
            pi = Real<R>::pos(a)[ki];
            si4 = Real<R>::prm(a)[ki];
            subI = (int) si4.z;             // subsets travel as exact small reals, -1 = padding lane
            iValid = subI >= 0;
            origI = a.sortedAtom[ki];
            cI = a.blockCenter[O>>2];
            oc = a.octetCenter[O];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int sl = sliceOf(max(subI, 0), s);
                t.lamC[s] = s < numSub ? (R) a.sp->lambdaC[sl] : (R) 0;
                t.lamV[s] = s < numSub ? (R) a.sp->lambdaV[sl] : (R) 0;
            }
            if (TWO) { dLamC = t.lamC[1]-t.lamC[0]; dLamV = t.lamV[1]-t.lamV[0]; }
        }
        perPair = PBC == PBC_SHIFT && (octCur&0x40000000) != 0;     // octet too wide for one image per j atom
        // a padding lane can never be "in range": its bounds are negative
        myCutHi = iValid ? (perPair ? c.cutoffHiWide : bandHi) : (R) -1;
        myCutLo = iValid ? (perPair ? c.cutoffLoWide : bandLo) : (R) -1;
        // ---- fix-up: the staged j atom of this lane moves into the octet's block frame
        // A CLEAN chunk -- no exclusion bit, no padding entry (mask 0xff), every j atom in the same subset -- takes the lean
        // copy of the pair loop: no mask loads, one lambda pair and one energy pair for the whole chunk (most chunks are)
        bool clean = false;
        R subJ0 = 0;
        {
            vec4 pj = St::ld4(sb+lane*St::STRIDE);
            const float4 cJ = ldSharedF4(sb+St::CTR+lane*16);
            fixupShift<R>(pj, cJ, cI, oc, PBC == PBC_SHIFT && !perPair, wb);
            St::st4(sb+lane*St::STRIDE, pj);
            if (CLEANABLE) {
                const unsigned mask = ldSharedU32(entBase+(unsigned) (ch&3)*256u+4u);
                const R subJ = *(const R*) (&sStage[buf][0]+St::PRM+lane*St::STRIDE+2*sizeof(R));
                subJ0 = __shfl_sync(full, subJ, 0);
                clean = !__any_sync(full, mask != 0u || subJ != subJ0);
            }
        }
        __syncwarp();
        R fix = 0, fiy = 0, fiz = 0, fjx = 0, fjy = 0, fjz = 0;
        R eCt = 0, eVt = 0, eC1 = 0, eV1 = 0;       // TWO: totals and the j-subset-1 parts
        unsigned bandMask = 0;
#pragma unroll
        for (int s = 0; s < NS; s++)
            t.eC[s] = t.eV[s] = 0;
        const unsigned own = sb+ownOff;
        // ---- 8 steps of 32 pairs: straight-line code (the rare per-pair-wrapping octets take their own copy) ----
        auto steps = [&](auto perPairTag, auto cleanTag) {
        constexpr bool PERPAIR = decltype(perPairTag)::value;
        constexpr bool CLEAN = decltype(cleanTag)::value;
        // CLEAN: the lambdas of the chunk's one j subset
        R lcC = t.lamC[0], lvC = t.lamV[0];
        if (CLEAN && NSUB > 1) {
#pragma unroll
            for (int u = 1; u < NS; u++) {
                lcC = subJ0 == (R) u ? t.lamC[u] : lcC;
                lvC = subJ0 == (R) u ? t.lamV[u] : lvC;
            }
        }
#pragma unroll STEP_UNROLL
        for (int s = 0; s < 8; s++) {
            const unsigned ad = own^(unsigned) (s*St::STRIDE);
            const unsigned aux = St::STRIDE == St::AUXS ? ad+St::AUX : sb+St::AUX+(unsigned) (lane^s)*St::AUXS;
            const vec4 q = St::ld4(ad);
            const unsigned mj = CLEAN ? 0u : ldSharedU32(aux+4);
            R dx = pi.x-q.x, dy = pi.y-q.y, dz = pi.z-q.z;
            if (PBC == PBC_SHIFT && PERPAIR) {
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
            if (!BRANCHY || in) {
                const vec4 prm = St::ld4(ad+St::PRM);
                R fc, fv, ec, ev, invR2;
                // masked lanes compute on r^2 = 1 (self pairs have r = 0, excluded neighbours overflow r^-12)
                pairMath<R, KIND, SWITCH>(c, BRANCHY ? r2 : (in ? r2 : (R) (SELF_MASKING ? 1e10 : 1)), pi.w*q.w, si4.x, si4.y, si4.w, prm.x, prm.y, prm.w, fc, fv, ec, ev, invR2);
                R lc, lv;
                if (NSUB == 1 || CLEAN) {
                    lc = lcC; lv = lvC;
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
                R f = (lv*fv+lc*fc)*invR2;
                if (!BRANCHY && !SELF_MASKING) {
                    f = in ? f : (R) 0;
                    ec = in ? ec : (R) 0;
                    ev = in ? ev : (R) 0;
                }
                fix += f*dx; fiy += f*dy; fiz += f*dz;
                fjx -= f*dx; fjy -= f*dy; fjz -= f*dz;
                if (ENERGY) {
                    if (CLEAN) {
                        eCt += ec; eVt += ev;
                    }
                    else if (TWO) {
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
                if (DUMP && in) {
                    const int origJ = a.sortedAtom[(int) ldSharedU32(aux)];
                    unsigned long long idx = atomicAdd(a.pairDumpCount, 1ull);
                    if ((long long) idx < a.pairDumpCapacity)
                        a.pairDump[idx] = make_int2(min(origI, origJ), max(origI, origJ));
                }
            }
            // the accumulators follow their slot: to the lane that takes it in the next step (home after the last)
            const int hop = s == 7 ? 7 : s^(s+1);
            fjx = __shfl_xor_sync(full, fjx, hop);
            fjy = __shfl_xor_sync(full, fjy, hop);
            fjz = __shfl_xor_sync(full, fjz, hop);
        }
        };
        if (PBC == PBC_SHIFT && perPair)
            steps(std::true_type(), std::false_type());     // rare (octets wider than half the box minus the cutoff): one copy
        else if (CLEANABLE && clean)
            steps(std::false_type(), std::integral_constant<bool, CLEANABLE>());
        else
            steps(std::false_type(), std::false_type());
        if (ENERGY && CLEANABLE && clean && !(PBC == PBC_SHIFT && perPair)) {
            // the chunk's energies belong to its one j subset
            if (TWO) { eC1 = subJ0*eCt; eV1 = subJ0*eVt; }
            else if (NSUB == 1) { t.eC[0] = eCt; t.eV[0] = eVt; }
            else if (NSUB > 2) {
#pragma unroll
                for (int u = 0; u < NS; u++) {
                    t.eC[u] = subJ0 == (R) u ? eCt : (R) 0;
                    t.eV[u] = subJ0 == (R) u ? eVt : (R) 0;
                }
            }
        }
        if (KIND != KIND_PLAIN && __any_sync(full, bandMask != 0u)) {
            if (bandMask) {
                DirectOut o;
                o.posd = a.posd; o.sortedAtom = a.sortedAtom; o.sp = a.sp; o.forceSorted = a.forceSorted; o.paddedAtoms = a.paddedAtoms;
                o.energyTable = a.energyBuf+(size_t) (blockIdx.x%ENERGY_REPLICAS)*2*SNB_MAX_SLICES;
                o.pairDump = a.pairDump; o.pairDumpCount = a.pairDumpCount; o.pairDumpCapacity = a.pairDumpCapacity;
                bandPairs<R, KIND, PBC, ENERGY, DUMP, SWITCH>(o, c, bandMask, lane, ki, origI, pi, si4, perPair ? 1 : 0, &sStage[buf][0]);
            }
        }
        const int jCur = (int) ldSharedU32(entBase+(unsigned) (ch&3)*256u);
        if (jCur >= 0) {
            addFixed(&a.forceSorted[jCur], Real<R>::fixed(fjx));
            addFixed(&a.forceSorted[a.paddedAtoms+jCur], Real<R>::fixed(fjy));
            addFixed(&a.forceSorted[2*a.paddedAtoms+jCur], Real<R>::fixed(fjz));
        }
        Fx += Real<R>::fixed(fix); Fy += Real<R>::fixed(fiy); Fz += Real<R>::fixed(fiz);
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
    if (curO >= 0)
        flushOctet();
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

template <int KIND, int PBC, int NSUB, bool SWITCH>
static void launch5(const DirectArgs& a, cudaStream_t st, int grid) {
    // every CTA resident at once (equal contiguous shares of the chunk list): more than two subsets take more registers
    if (NSUB > 2 || NSUB == 0)
        grid = grid/DIRECT_CTAS_PER_SM*DIRECT_WIDE_CTAS_PER_SM;
    if (a.doublePrecision) {
        // the double-precision mode is a correctness mode: one instantiation (energies always on)
        if (a.pairDump) k_direct<double, KIND, PBC, NSUB, true, true, SWITCH><<<grid/2, 32, 0, st>>>(a);
        else k_direct<double, KIND, PBC, NSUB, true, false, SWITCH><<<grid/2, 32, 0, st>>>(a);
        return;
    }
    if (a.pairDump)
        k_direct<float, KIND, PBC, NSUB, true, true, SWITCH><<<grid, 32, 0, st>>>(a);
    else if (a.needEnergy)
        k_direct<float, KIND, PBC, NSUB, true, false, SWITCH><<<grid, 32, 0, st>>>(a);
    else
        k_direct<float, KIND, PBC, NSUB, false, false, SWITCH><<<grid, 32, 0, st>>>(a);
}

template <int KIND, int PBC, int NSUB>
static void launch4(const DirectArgs& a, cudaStream_t st, int grid) {
    // the switching function exists for the cutoff and Ewald/PME kinds only (LJPME and NoCutoff have none)
    if ((KIND == KIND_RF || KIND == KIND_EWALD) && a.useSwitch)
        launch5<KIND, PBC, NSUB, (KIND == KIND_RF || KIND == KIND_EWALD)>(a, st, grid);
    else
        launch5<KIND, PBC, NSUB, false>(a, st, grid);
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
