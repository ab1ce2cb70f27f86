// Direct-space tile kernel (north-star subsystem 2).
//
// One warp per work item = one i block (32 sorted atoms, one per lane, kept in registers) against
// up to SNB_ITEM_CHUNKS chunks of 32 compacted j atoms.  A chunk's j atoms are gathered with
// coalesced float4 loads, shifted into the i block's frame and staged in shared memory; in step s
// lane l works on slot (l+s)&31, so every step touches 32 different j atoms, the j-force
// accumulators travel between lanes by warp shuffle and come home after 32 steps, and no
// reduction tree is needed.  Per-pair arithmetic is FP32 (rsqrt, ex2, polynomial erfc); the
// i forces are carried in FP64 across chunks, all forces leave as 2^32 fixed-point 64-bit atomics
// (order independent => bit-reproducible), slice energies are accumulated per j subset in FP32
// per chunk and FP64 beyond.
//
// Pair arithmetic follows the reference's Reference platform:
//   Ewald/PME/LJPME direct:  platforms/reference/src/ReferenceSlicedLJCoulombIxn.cpp:351-429
//   cutoff / no cutoff:      ReferenceSlicedLJCoulombIxn.cpp:553-611 (reaction field :64-65)
// (behavioural cross-reference on the GPU side: platforms/common/src/kernels/coulombLennardJones.cc:1-124).
#include "kernels.cuh"

#define DIRECT_WARPS 8

template <int KIND> void launchDirectKind(const DirectArgs& a, cudaStream_t st, int grid);

template <int NSUB> struct SubsetTable {
    // lambda and energy tables indexed by the j atom's subset (the i subset is fixed per lane)
    float lamC[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS], lamV[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS];
    float eC[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS], eV[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS];
    double EC[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS], EV[NSUB > 0 ? NSUB : SNB_MAX_SUBSETS];
};

template <int KIND, int PBC, int NSUB, bool ENERGY>
__global__ void __launch_bounds__(DIRECT_WARPS*32) k_direct(DirectArgs a) {
    constexpr int NS = NSUB > 0 ? NSUB : SNB_MAX_SUBSETS;
    __shared__ float4 sPos[DIRECT_WARPS][32];
    __shared__ float4 sPrm[DIRECT_WARPS][32];
    __shared__ int2 sEnt[DIRECT_WARPS][32];
    __shared__ double sEnergy[2*SNB_MAX_SLICES];
    const int lane = threadIdx.x&31, w = threadIdx.x>>5;
    const unsigned full = 0xffffffffu;
    const DirectParams& p = a.p;
    const int numSub = NSUB > 0 ? NSUB : a.numSubsets;
    const bool SWITCH = a.useSwitch != 0, DUMP = a.pairDump != nullptr;
    const int numItems = a.counters[1];
    if (ENERGY) {
        for (int k = threadIdx.x; k < 2*SNB_MAX_SLICES; k += blockDim.x)
            sEnergy[k] = 0.0;
        __syncthreads();
    }
    // per-lane energy totals keyed by (own subset at the time, j subset) are flushed per item
    for (;;) {
        int item = 0;
        if (lane == 0)
            item = atomicAdd(&a.counters[3], 1);
        item = __shfl_sync(full, item, 0);
        if (item >= numItems)
            break;
        const int4 wi = a.workItems[item];
        const int I = wi.x;
        const int ki = I*32+lane;
        const float4 pi = a.posqLocal[ki];
        const float4 si4 = a.sparams[ki];
        const int subI = __float_as_int(si4.z);
        const bool iValid = subI >= 0;
        const int origI = a.sortedAtom[ki];
        const double cIx = a.blockCenter[3*I], cIy = a.blockCenter[3*I+1], cIz = a.blockCenter[3*I+2];
        SubsetTable<NSUB> t;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            int sl = sliceOf(max(subI, 0), s);
            t.lamC[s] = s < numSub ? p.lambdaC[sl] : 0.f;
            t.lamV[s] = s < numSub ? p.lambdaV[sl] : 0.f;
            t.EC[s] = t.EV[s] = 0.0;
        }
        double Fx = 0, Fy = 0, Fz = 0;
        for (int c = 0; c < wi.z; c++) {
            // ---- gather and stage the chunk's j atoms -------------------------------------
            const int2 ent = a.jList[(size_t) (wi.y+c)*32+lane];
            float4 pj = make_float4(0.f, 0.f, 0.f, 0.f), sj4 = make_float4(0.f, 0.f, __int_as_float(0), 0.f);
            if (ent.x >= 0) {
                pj = a.posqLocal[ent.x];
                sj4 = a.sparams[ent.x];
                const int J = ent.x>>5;
                float ox = (float) (a.blockCenter[3*J]-cIx), oy = (float) (a.blockCenter[3*J+1]-cIy), oz = (float) (a.blockCenter[3*J+2]-cIz);
                pj.x += ox; pj.y += oy; pj.z += oz;
                if (PBC == PBC_SHIFT) {
                    // one image per j atom, chosen relative to the i block's centre (valid while
                    // halfBox - cutoff exceeds the block's half extent; validated in k_blockBounds' caller)
                    pj.x -= p.box.ax*rintf(pj.x*p.box.invAx);
                    pj.y -= p.box.by*rintf(pj.y*p.box.invBy);
                    pj.z -= p.box.cz*rintf(pj.z*p.box.invCz);
                }
            }
            __syncwarp();
            sPos[w][lane] = pj;
            sPrm[w][lane] = sj4;
            sEnt[w][lane] = ent;
            __syncwarp();
            float fix = 0.f, fiy = 0.f, fiz = 0.f, fjx = 0.f, fjy = 0.f, fjz = 0.f;
#pragma unroll
            for (int s = 0; s < NS; s++)
                t.eC[s] = t.eV[s] = 0.f;
            // ---- 32 x 32 pair steps ------------------------------------------------------
#pragma unroll 4
            for (int s = 0; s < 32; s++) {
                const int slot = (lane+s)&31;
                const float4 q = sPos[w][slot];
                const unsigned mj = (unsigned) sEnt[w][slot].y;
                float dx = pi.x-q.x, dy = pi.y-q.y, dz = pi.z-q.z;
                if (PBC == PBC_PAIR_ORTHO) {
                    dx -= p.box.ax*rintf(dx*p.box.invAx);
                    dy -= p.box.by*rintf(dy*p.box.invBy);
                    dz -= p.box.cz*rintf(dz*p.box.invCz);
                }
                else if (PBC == PBC_PAIR_TRICLINIC) {
                    float k = rintf(dz*p.box.invCz);
                    dx -= k*p.box.cx; dy -= k*p.box.cy; dz -= k*p.box.cz;
                    k = rintf(dy*p.box.invBy);
                    dx -= k*p.box.bx; dy -= k*p.box.by;
                    k = rintf(dx*p.box.invAx);
                    dx -= k*p.box.ax;
                }
                const float r2 = dx*dx+dy*dy+dz*dz;
                bool in = iValid && !((mj>>lane)&1u);
                if (KIND != KIND_PLAIN) {
                    in = in && r2 < p.cutoff2;
                    // Borderline pairs are decided exactly as the CPU oracle decides them: in double,
                    // from the untouched input coordinates (bit-exact pair set).
                    if (iValid && !((mj>>lane)&1u) && fabsf(r2-p.cutoff2) < p.bandTol) {
                        const int origJ = a.sortedAtom[sEnt[w][slot].x];
                        double d[3];
                        deltaExact(a.posd+3*(size_t) min(origI, origJ), a.posd+3*(size_t) max(origI, origJ), p.boxd, p.periodic != 0, d);
                        in = !(dot3Exact(d) > p.cutoff2d);
                    }
                }
                if (in) {
                    const float4 prm = sPrm[w][slot];
                    const int subJ = __float_as_int(prm.z);
                    const float invR = fastRsqrt(r2);
                    const float invR2 = invR*invR;
                    const float r = r2*invR;
                    const float qq = pi.w*q.w;                  // charges carry sqrt(ONE_4PI_EPS0)
                    float fc, ec;
                    if (KIND == KIND_EWALD || KIND == KIND_LJPME) {
                        const float ar = p.alpha*r;
                        const float ex = fastExp2(-1.4426950408889634f*(p.alpha2*r2));
                        const float erfcAR = ex*erfcxPoly(ar);
                        ec = qq*invR*erfcAR;
                        fc = qq*invR*invR2*fmaf(p.twoAlphaOverSqrtPi*r, ex, erfcAR);
                    }
                    else if (KIND == KIND_RF) {
                        ec = qq*(invR+p.krf*r2-p.crf);
                        fc = qq*invR2*(invR-2.0f*p.krf*r2);
                    }
                    else {
                        ec = qq*invR;
                        fc = ec*invR2;
                    }
                    const float sig = si4.x+prm.x;
                    const float eps = si4.y*prm.y;
                    float s2 = sig*invR;
                    s2 *= s2;
                    const float s6 = s2*s2*s2;
                    float fv = eps*(12.0f*s6-6.0f)*s6*invR2;
                    float ev = eps*(s6-1.0f)*s6;
                    if (KIND == KIND_LJPME) {
                        const float c6 = si4.w*prm.w;
                        const float dar2 = p.dalpha2*r2, dar4 = dar2*dar2;
                        const float exd = fastExp2(-1.4426950408889634f*dar2);
                        const float invR6 = invR2*invR2*invR2;
                        const float pe = 1.0f+dar2+0.5f*dar4;
                        ev += c6*invR6*(1.0f-exd*pe);
                        fv += 6.0f*c6*invR6*invR2*(1.0f-exd*(pe+dar4*dar2*(1.0f/6.0f)));
                        float sc2 = sig*sig;
                        const float sc6 = sc2*sc2*sc2*p.invCut6;
                        ev += eps*(1.0f-sc6)*sc6-c6*p.invCut6*p.dispShiftExp;
                    }
                    if (SWITCH) {
                        if (r > p.switchDist) {
                            const float x = (r-p.switchDist)*p.invSwitchWidth;
                            const float sw = 1.0f+x*x*x*(-10.0f+x*(15.0f-x*6.0f));
                            const float dsw = x*x*(-30.0f+x*(60.0f-x*30.0f))*p.invSwitchWidth;
                            fv = sw*fv-ev*dsw*invR;
                            ev *= sw;
                        }
                    }
                    float lc, lv;
                    if (NSUB == 1) {
                        lc = t.lamC[0]; lv = t.lamV[0];
                    }
                    else if (NSUB > 0) {
                        lc = t.lamC[0]; lv = t.lamV[0];
#pragma unroll
                        for (int u = 1; u < NS; u++) {
                            lc = subJ == u ? t.lamC[u] : lc;
                            lv = subJ == u ? t.lamV[u] : lv;
                        }
                    }
                    else {
                        const int sl = sliceOf(subI, subJ);
                        lc = p.lambdaC[sl]; lv = p.lambdaV[sl];
                    }
                    const float f = lv*fv+lc*fc;
                    fix = fmaf(f, dx, fix); fiy = fmaf(f, dy, fiy); fiz = fmaf(f, dz, fiz);
                    fjx = fmaf(-f, dx, fjx); fjy = fmaf(-f, dy, fjy); fjz = fmaf(-f, dz, fjz);
                    if (ENERGY) {
                        if (NSUB > 0) {
#pragma unroll
                            for (int u = 0; u < NS; u++) {
                                t.eC[u] += subJ == u ? ec : 0.f;
                                t.eV[u] += subJ == u ? ev : 0.f;
                            }
                        }
                        else {
                            // many subsets: straight to the block's shared table (rare configuration)
                            atomicAdd(&sEnergy[2*sliceOf(subI, subJ)], (double) ec);
                            atomicAdd(&sEnergy[2*sliceOf(subI, subJ)+1], (double) ev);
                        }
                    }
                    if (DUMP) {
                        const int origJ = a.sortedAtom[sEnt[w][slot].x];
                        unsigned long long idx = atomicAdd(a.pairDumpCount, 1ull);
                        if ((long long) idx < a.pairDumpCapacity)
                            a.pairDump[idx] = make_int2(min(origI, origJ), max(origI, origJ));
                    }
                }
                // the j accumulators move one lane down; after 32 steps they are home
                fjx = __shfl_sync(full, fjx, (lane+1)&31);
                fjy = __shfl_sync(full, fjy, (lane+1)&31);
                fjz = __shfl_sync(full, fjz, (lane+1)&31);
            }
            if (ent.x >= 0) {
                addFixed(&a.forceSorted[ent.x], toFixed(fjx));
                addFixed(&a.forceSorted[a.paddedAtoms+ent.x], toFixed(fjy));
                addFixed(&a.forceSorted[2*a.paddedAtoms+ent.x], toFixed(fjz));
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
        if (iValid) {
            addFixed(&a.forceSorted[ki], toFixedD(Fx));
            addFixed(&a.forceSorted[a.paddedAtoms+ki], toFixedD(Fy));
            addFixed(&a.forceSorted[2*a.paddedAtoms+ki], toFixedD(Fz));
        }
        if (ENERGY && NSUB > 0 && iValid) {
#pragma unroll
            for (int u = 0; u < NS; u++) {
                if (t.EC[u] != 0.0) atomicAdd(&sEnergy[2*sliceOf(subI, u)], t.EC[u]);
                if (t.EV[u] != 0.0) atomicAdd(&sEnergy[2*sliceOf(subI, u)+1], t.EV[u]);
            }
        }
    }
    if (ENERGY) {
        __syncthreads();
        const int L = numSub*(numSub+1)/2;
        for (int k = threadIdx.x; k < 2*L; k += blockDim.x)
            if (sEnergy[k] != 0.0)
                atomicAdd(&a.energyBuf[k], sEnergy[k]);
    }
}

template <int KIND, int PBC, int NSUB>
static void launch4(const DirectArgs& a, cudaStream_t st, int grid) {
    if (a.needEnergy || a.pairDump)
        k_direct<KIND, PBC, NSUB, true><<<grid, DIRECT_WARPS*32, 0, st>>>(a);
    else
        k_direct<KIND, PBC, NSUB, false><<<grid, DIRECT_WARPS*32, 0, st>>>(a);
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
            if (KIND != KIND_PLAIN) launch3<KIND, PBC_SHIFT>(a, st, grid);
            break;
        case PBC_PAIR_ORTHO:
            if (KIND != KIND_PLAIN) launch3<KIND, PBC_PAIR_ORTHO>(a, st, grid);
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
