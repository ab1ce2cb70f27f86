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
    static __device__ __forceinline__ float rsqrt(float r2) {
        // MUFU.RSQ (2 ulp) + one Newton step: the LJ force goes as invR^14, so the seed's error
        // would otherwise dominate the pair error
        float y = fastRsqrt(r2);
        return y*fmaf(-0.5f*r2, y*y, 1.5f);
    }
    static __device__ __forceinline__ float expNeg(float x) { return fastExp2(-1.4426950408889634f*x); }    // exp(-x)
    static __device__ __forceinline__ void erfcTerms(float ar, float a2r2, float& ex, float& erfcAR) {
        ex = expNeg(a2r2);
        erfcAR = ex*erfcxPoly(ar);
    }
    static __device__ __forceinline__ float rint(float x) { return rintf(x); }
    static __device__ __forceinline__ long long fixed(float f) { return toFixed(f); }
};
template <> struct Real<double> {
    typedef double4 vec4;
    static __device__ __forceinline__ const double4* pos(const DirectArgs& a) { return a.posqLocalD; }
    static __device__ __forceinline__ const double4* prm(const DirectArgs& a) { return a.sparamsD; }
    static __device__ __forceinline__ double rsqrt(double r2) { return 1.0/sqrt(r2); }
    static __device__ __forceinline__ double expNeg(double x) { return exp(-x); }
    static __device__ __forceinline__ void erfcTerms(double ar, double a2r2, double& ex, double& erfcAR) {
        ex = exp(-a2r2);
        erfcAR = erfc(ar);
    }
    static __device__ __forceinline__ double rint(double x) { return ::rint(x); }
    static __device__ __forceinline__ long long fixed(double f) { return toFixedD(f); }
};

// Borderline pairs (|r^2 - rc^2| within the single-precision uncertainty) are decided exactly as the
// CPU oracle decides them: in double, from the untouched input coordinates, lower index first.  Kept
// out of line: it runs for ~1e-5 of the pairs and would otherwise bloat the hot loop.
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

template <typename R, int KIND, int PBC, int NSUB, bool ENERGY>
__global__ void __launch_bounds__(32, sizeof(R) == 4 ? (NSUB > 2 || NSUB == 0 ? DIRECT_CTAS_PER_SM/2 : DIRECT_CTAS_PER_SM) : 8) k_direct(DirectArgs a) {
    constexpr int NS = NSUB > 0 ? NSUB : SNB_MAX_SUBSETS;
    typedef typename Real<R>::vec4 vec4;
    // one warp per CTA: the staging area is addressed by the slot alone
    __shared__ vec4 sPos[32];
    __shared__ vec4 sPrm[32];
    __shared__ unsigned sMask[32];
    __shared__ int sJ[32];
    const int lane = threadIdx.x;
    const int copy = lane>>3, k = lane&7;
    const unsigned full = 0xffffffffu;
    const DirectParams& p = a.p;
    const int numSub = NSUB > 0 ? NSUB : a.numSubsets;
    const bool SWITCH = a.useSwitch != 0, DUMP = a.pairDump != nullptr;
    const int numItems = a.counters[1];
    const R cutoffHi = (R) (p.cutoff2d+p.bandTol), cutoffLo = (R) (p.cutoff2d-p.bandTol);
    const R alpha = (R) p.alphad, alpha2 = alpha*alpha, twoAlphaOverSqrtPi = (R) (2.0*p.alphad/1.772453850905516);
    const R krf = (R) p.krfd, crf = (R) p.crfd;
    const R switchDist = (R) p.switchDistd, invSwitchWidth = (R) p.invSwitchWidthd;
    const R dalpha2 = (R) (p.dalphad*p.dalphad), dispShiftExp = (R) p.dispShiftExpd, invCut6 = (R) p.invCut6d;
    const BoxD boxd = a.sp->exact.boxd;
    const R bAx = (R) boxd.ax, bBy = (R) boxd.by, bCz = (R) boxd.cz;
    const R iAx = (R) (1.0/boxd.ax), iBy = (R) (1.0/boxd.by), iCz = (R) (1.0/boxd.cz);
    const unsigned kBit = 1u<<k;
    const int srcLane = (copy<<3)|((k+1)&7);       // the j accumulators move one lane down inside the copy
    // slice energies of this thread over all its work items, keyed by (own subset, j subset); reduced
    // over the warp once, at the end of the kernel
    double accC[NS > 4 ? 1 : NS][NS > 4 ? 1 : NS], accV[NS > 4 ? 1 : NS][NS > 4 ? 1 : NS];
    if (ENERGY && NSUB > 0) {
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int u = 0; u < NS; u++)
                accC[s][u] = accV[s][u] = 0.0;
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
        // a padding lane can never be "in range": its upper bound is negative
        const R myCutHi = iValid ? cutoffHi : (R) -1;
        SubsetTable<R, NSUB> t;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            int sl = sliceOf(max(subI, 0), s);
            t.lamC[s] = s < numSub ? (R) a.sp->lambdaC[sl] : (R) 0;
            t.lamV[s] = s < numSub ? (R) a.sp->lambdaV[sl] : (R) 0;
            t.EC[s] = t.EV[s] = 0.0;
        }
        double Fx = 0, Fy = 0, Fz = 0;
        for (int c = 0; c < wi.z; c++) {
            // ---- gather and stage the chunk's j atoms -------------------------------------
            const int2 ent = a.jList[(size_t) (wi.y+c)*32+lane];
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
                    // halfBox - cutoff exceeds the octet's half extent; k_compactList flags the rest)
                    pj.x -= bAx*Real<R>::rint((pj.x-(R) oc.x)*iAx);
                    pj.y -= bBy*Real<R>::rint((pj.y-(R) oc.y)*iBy);
                    pj.z -= bCz*Real<R>::rint((pj.z-(R) oc.z)*iCz);
                }
            }
            __syncwarp();
            sPos[lane] = pj;
            sPrm[lane] = sj4;
            sMask[lane] = (unsigned) ent.y;
            sJ[lane] = ent.x;
            __syncwarp();
            R fix = 0, fiy = 0, fiz = 0, fjx = 0, fjy = 0, fjz = 0;
#pragma unroll
            for (int s = 0; s < NS; s++)
                t.eC[s] = t.eV[s] = 0;
            // ---- 8 steps of 32 pairs --------------------------------------------------------
#pragma unroll 2
            for (int s = 0; s < 8; s++) {
                const int slot = (copy<<3)|((k+s)&7);
                const vec4 q = sPos[slot];
                const unsigned mj = sMask[slot];
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
                // one compare on the hot path: r2 below cutoff2 + band; pairs inside the band (about 1e-5
                // of them) are then decided exactly, as the CPU oracle decides them
                bool in = !(mj&kBit);
                if (KIND != KIND_PLAIN) {
                    in = in && r2 < myCutHi;
                    if (in && r2 > cutoffLo)
                        in = exactInRange(a.posd, origI, a.sortedAtom[sJ[slot]], &a.sp->exact);
                }
                else
                    in = in && iValid;
                if (in) {
                    const vec4 prm = sPrm[slot];
                    const R invR = Real<R>::rsqrt(r2);
                    const R invR2 = invR*invR;
                    const R r = r2*invR;
                    const R qq = pi.w*q.w;                      // charges carry sqrt(ONE_4PI_EPS0)
                    R fc, ec;
                    if (KIND == KIND_EWALD || KIND == KIND_LJPME) {
                        R ex, erfcAR;
                        Real<R>::erfcTerms(alpha*r, alpha2*r2, ex, erfcAR);
                        ec = qq*invR*erfcAR;
                        fc = qq*invR*invR2*(twoAlphaOverSqrtPi*r*ex+erfcAR);
                    }
                    else if (KIND == KIND_RF) {
                        ec = qq*(invR+krf*r2-crf);
                        fc = qq*invR2*(invR-2*krf*r2);
                    }
                    else {
                        ec = qq*invR;
                        fc = ec*invR2;
                    }
                    const R sig = si4.x+prm.x;
                    const R eps = si4.y*prm.y;
                    R s2 = sig*invR;
                    s2 *= s2;
                    const R s6 = s2*s2*s2;
                    R fv = eps*(12*s6-6)*s6*invR2;
                    R ev = eps*(s6-1)*s6;
                    if (KIND == KIND_LJPME) {
                        const R c6 = si4.w*prm.w;
                        const R dar2 = dalpha2*r2, dar4 = dar2*dar2;
                        const R exd = Real<R>::expNeg(dar2);
                        const R invR6 = invR2*invR2*invR2;
                        const R pe = 1+dar2+(R) 0.5*dar4;
                        ev += c6*invR6*(1-exd*pe);
                        fv += 6*c6*invR6*invR2*(1-exd*(pe+dar4*dar2*((R) 1/(R) 6)));
                        R sc2 = sig*sig;
                        const R sc6 = sc2*sc2*sc2*invCut6;
                        ev += eps*(1-sc6)*sc6-c6*invCut6*dispShiftExp;
                    }
                    if (SWITCH) {
                        if (r > switchDist) {
                            const R x = (r-switchDist)*invSwitchWidth;
                            const R sw = 1+x*x*x*(-10+x*(15-x*6));
                            const R dsw = x*x*(-30+x*(60-x*30))*invSwitchWidth;
                            fv = sw*fv-ev*dsw*invR;
                            ev *= sw;
                        }
                    }
                    R lc, lv;
                    if (NSUB == 1) {
                        lc = t.lamC[0]; lv = t.lamV[0];
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
                    const R f = lv*fv+lc*fc;
                    fix += f*dx; fiy += f*dy; fiz += f*dz;
                    fjx -= f*dx; fjy -= f*dy; fjz -= f*dz;
                    if (ENERGY) {
                        if (NSUB > 0) {
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
                        const int origJ = a.sortedAtom[sJ[slot]];
                        unsigned long long idx = atomicAdd(a.pairDumpCount, 1ull);
                        if ((long long) idx < a.pairDumpCapacity)
                            a.pairDump[idx] = make_int2(min(origI, origJ), max(origI, origJ));
                    }
                }
                fjx = __shfl_sync(full, fjx, srcLane);
                fjy = __shfl_sync(full, fjy, srcLane);
                fjz = __shfl_sync(full, fjz, srcLane);
            }
            if (ent.x >= 0) {
                addFixed(&a.forceSorted[ent.x], Real<R>::fixed(fjx));
                addFixed(&a.forceSorted[a.paddedAtoms+ent.x], Real<R>::fixed(fjy));
                addFixed(&a.forceSorted[2*a.paddedAtoms+ent.x], Real<R>::fixed(fjz));
            }
            Fx += fix; Fy += fiy; Fz += fiz;
            if (ENERGY && NSUB > 0) {
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
        if (ENERGY && NSUB > 0) {
#pragma unroll
            for (int s = 0; s < NS; s++)
#pragma unroll
                for (int u = 0; u < NS; u++) {
                    accC[s][u] += subI == s ? t.EC[u] : 0.0;
                    accV[s][u] += subI == s ? t.EV[u] : 0.0;
                }
        }
    }
    if (ENERGY && NSUB > 0) {
        double* table = a.energyBuf+(size_t) (blockIdx.x%ENERGY_REPLICAS)*2*SNB_MAX_SLICES;
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int u = 0; u < NS; u++) {
                double ecs = accC[s][u], evs = accV[s][u];
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
        k_direct<double, KIND, PBC, NSUB, true><<<grid/3, 32, 0, st>>>(a);
        return;
    }
    if (a.needEnergy || a.pairDump)
        k_direct<float, KIND, PBC, NSUB, true><<<grid, 32, 0, st>>>(a);
    else
        k_direct<float, KIND, PBC, NSUB, false><<<grid, 32, 0, st>>>(a);
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
