// Direct-space tile kernel, packed-FP32 version for the headline path: PME/Ewald direct space, orthorhombic box,
// one or two subsets, no switching function, mixed precision.  Everything else runs k_direct (direct_impl.cuh).
//
// Same tile list, same chunk pipeline (list entry two chunks ahead, cp.async gather one chunk ahead), same 2^32
// fixed-point accumulation as k_direct.  What changes is the pair loop: sm_100 has two-wide FP32 instructions
// (FFMA2 / FMUL2 / FADD2: two independent IEEE operations per lane per issue slot; measured on B200 the FMA pipe
// delivers the same 70 TFLOP/s either way, but the packed form needs half the issue slots, and k_direct is
// issue-bound: tools/ubench/fp32peak.cu).  So every lane works on TWO j atoms at a time:
//   * after the fix-up pass the 32 staged j atoms are re-laid out as 16 slot pairs, structure-of-arrays
//     within the pair -- (x0,x1,y0,y1) (z0,z1,q0,q1) (sig0,sig1,eps0,eps1) (sub0,sub1,mask0,mask1) -- so one
//     LDS.128 hands the packed instructions their operands in aligned register pairs;
//   * a chunk is 4 double-steps: lane (copy c, atom k) takes the slot pair 4c + ((k>>1)^d) in double-step d,
//     at the address own^(16 d); the i atom's coordinates / parameters and all constants enter as broadcast
//     operands (FFMA2 takes a scalar register or an immediate for both halves);
//   * the j-force accumulators of a slot pair travel with it (butterfly shuffles over lane masks 2, 6, 2, 6);
//     lanes k and k^1 hold the two partial sums of the same pair and swap halves once per chunk;
//   * only MUFU (rsqrt, ex2, rcp) and the cutoff / exclusion predicates remain scalar.
// Pairs inside the band around the cutoff are noted by a predicate only and settled exactly afterwards by
// bandRescan (direct_impl.cuh), which recomputes r^2 with the very same roundings: pairDelta pins them with explicit
// intrinsics, and the packed loop follows the same sequence (x - q == fma(q, -1, x) exactly).
//
// Pair arithmetic: platforms/reference/src/ReferenceSlicedLJCoulombIxn.cpp:351-429.
#include "direct_impl.cuh"

namespace {

__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ void stSharedF32(unsigned ad, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(ad), "f"(v) : "memory"); }
__device__ __forceinline__ void stSharedU32(unsigned ad, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(ad), "r"(v) : "memory"); }

#ifndef PACKED_CTAS_PER_SM
#define PACKED_CTAS_PER_SM 12       /* one-warp CTAs of the packed kernel resident per SM: 168 registers, no spills (16 -> 128 registers spill) */
#endif

// pair-layout planes of the 16 slot pairs (16 B per pair and plane)
enum { PL_XY = 0, PL_ZQ = 256, PL_SE = 512, PL_SM = 768, PL_BYTES = 1024 };


// The pairs of one chunk that fell inside the band around the cutoff: the pair loop only notes that there is one; here the
// lane finds them again in the pair planes (same r^2, bit for bit: pairDelta), decides each exactly, evaluates it and adds it
// with atomics, so that the pair loop itself carries neither a call nor a per-step flag.  (bandRescan of direct_impl.cuh, reading
// the packed kernel's own layout: slot = pair 2P+h -> word h of the pair's (x0,x1,y0,y1) (z0,z1,q0,q1) (s0,s1,e0,e1) (sub0,sub1,m0,m1).)
template <bool ENERGY, bool DUMP>
static __device__ __noinline__ void bandRescanPlanes(DirectOut o, DirectConsts<float> c, int lane, int ki, int origI, float cutLo, float cutHi,
                                                     float4 pi, float4 si4, int perPair, const unsigned char* planes, const int2* entries) {
    const int copy = lane>>3, k = lane&7;
    const WrapBox<float> wb = makeWrapBox<float>(o.sp->exact.boxd);
    for (int s = 0; s < 8; s++) {
        const int slot = (copy<<3)|(k^s);
        const float* P = (const float*) (planes+(slot>>1)*16)+(slot&1);
        const float4 q = make_float4(P[PL_XY/4], P[PL_XY/4+2], P[PL_ZQ/4], P[PL_ZQ/4+2]);
        const unsigned mask = __float_as_uint(P[PL_SM/4+2]);
        float dx, dy, dz, r2;
        if (perPair) r2 = pairDelta<1>(pi, q, wb, dx, dy, dz);
        else r2 = pairDelta<0>(pi, q, wb, dx, dy, dz);
        if ((mask>>k)&1u || r2 < cutLo || !(r2 < cutHi))
            continue;
        const float sigJ = P[PL_SE/4], epsJ = P[PL_SE/4+2], subJ = P[PL_SM/4];
        const int j = entries[slot].x;
        const int origJ = o.sortedAtom[j];
        if (!exactInRange(o.posd, origI, origJ, &o.sp->exact))
            continue;
        float fc, fv, ec, ev, invR2;
        pairMath<float, KIND_EWALD, false>(c, r2, pi.w*q.w, si4.x, si4.y, si4.w, sigJ, epsJ, 0.f, fc, fv, ec, ev, invR2);
        const int slice = sliceOf((int) si4.z, (int) subJ);
        const float f = ((float) o.sp->lambdaV[slice]*fv+(float) o.sp->lambdaC[slice]*fc)*invR2;
        const long long fx = toFixed(f*dx), fy = toFixed(f*dy), fz = toFixed(f*dz);
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

template <int NSUB, bool ENERGY, bool DUMP>
__global__ void __launch_bounds__(32, PACKED_CTAS_PER_SM) k_directPacked(DirectArgs a) {
    static_assert(NSUB == 1 || NSUB == 2, "packed pair loop: one or two subsets");
    typedef Stage<float> St;
    __shared__ __align__(256) unsigned char sStage[2][St::BYTES];
    __shared__ __align__(256) unsigned char sPairs[2][PL_BYTES];       // pair planes of the chunk in the loop and of the next one
    __shared__ __align__(16) int2 sEnt[4][32];          // list entries of chunks ch .. ch+2 (ring of four)
    __shared__ int sOct[4];                             // and their octets
    __shared__ double sAccC[NSUB*NSUB][32], sAccV[NSUB*NSUB][32];
    constexpr unsigned BUF = sizeof(sStage[0]);
    const int lane = threadIdx.x;
    const int copy = lane>>3, k = lane&7;
    const unsigned full = 0xffffffffu;
    const DirectConsts<float>& c = a.cf;
    // a list that overflowed its buffers holds unwritten chunks: the host discards this pass and repeats it
    const int numChunks = __shfl_sync(full, listChunks(a), 0);
    const int per = (numChunks+(int) gridDim.x-1)/(int) gridDim.x;
    const int cBegin = min(numChunks, (int) blockIdx.x*per), cEnd = min(numChunks, cBegin+per);
    const WrapBox<float> wb = makeWrapBox<float>(a.sp->exact.boxd);
    // cutoff band: narrow unless some block of this evaluation is large (see k_direct)
    const bool narrowOk = !(__int_as_float(a.counters[SNB_COUNTER_MAX_EXTENT]) > SNB_NARROW_MAX_EXTENT);
    const float bandHi = narrowOk ? c.cutoffHi : c.cutoffHiWide, bandLo = narrowOk ? c.cutoffLo : c.cutoffLoWide;
    const unsigned kBit = opaque(1u<<k);
    const unsigned stageBase = opaque((unsigned) __cvta_generic_to_shared(&sStage[0][0]));
    const unsigned pairBase = opaque((unsigned) __cvta_generic_to_shared(&sPairs[0][0]));
    const unsigned entBase = opaque((unsigned) __cvta_generic_to_shared(&sEnt[0][0])+(unsigned) lane*8u);
    const unsigned octBase = opaque((unsigned) __cvta_generic_to_shared(&sOct[0]));
    const unsigned ownPair = opaque(pairBase+(unsigned) (lane>>1)*16u);                 // the slot pair of double-step 0
    const unsigned ownHalf = opaque(pairBase+(unsigned) (lane>>1)*16u+(unsigned) (lane&1)*4u);   // where this lane's own j atom goes
    // the two components a lane stores into a plane go out in two instructions; lanes of the upper eight pairs store them
    // in the other order, so that each instruction's 32 words fall into 32 different banks
    const bool swapStores = lane >= 16;
    const unsigned st1 = opaque(ownHalf+(swapStores ? 8u : 0u)), st2 = opaque(ownHalf+(swapStores ? 0u : 8u));
    const bool odd = (lane&1) != 0;
    if (ENERGY) {
#pragma unroll
        for (int s = 0; s < NSUB*NSUB; s++)
            sAccC[s][lane] = sAccV[s][lane] = 0.0;
    }

    // Everything the steady-state loop reads from global memory arrives through cp.async: a plain load would hold a
    // register scoreboard for a DRAM round trip, and with six scoreboards per warp some unrelated instruction of the
    // pair loop ends up waiting on it (ncu: ~25% of the warp time in long-scoreboard stalls with plain prefetch loads).
    //   fetchEntries(c): list entry of every lane + the octet of chunk c  -> ring slot c&3      (two chunks ahead)
    //   gather(c):       the 32 j atoms of chunk c (entries already in the ring) -> stage c&1    (one chunk ahead)
    auto fetchEntries = [&](int chunk) {
        cpAsync8(entBase+(unsigned) (chunk&3)*256u, a.jList+(size_t) chunk*32+lane);
        if (lane == 0)
            cpAsync4(octBase+(unsigned) (chunk&3)*4u, a.chunkOctet+chunk);
    };
    auto gather = [&](int chunk) {
        const int j = max((int) ldSharedU32(entBase+(unsigned) (chunk&3)*256u), 0);    // padding entries copy atom 0: never in range (mask 0xff)
        const unsigned dst = stageBase+(unsigned) (chunk&1)*BUF;
        cpAsync16(dst+lane*St::STRIDE, a.posqLocal+j);
        cpAsync16(dst+St::PRM+lane*St::STRIDE, a.sparams+j);
        cpAsync16(dst+St::CTR+lane*16, a.blockCenter+(j>>5));
    };

    // ---- the octet in flight
    int curO = -1, ki = 0, origI = -1, subI = -1;
    bool iValid = false;
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), si4 = pi;
    float4 cI = pi, oc = pi;
    float myCutHi = -1.f, myCutLo = -1.f, myCutMid = -1.f, myBandTol = 0.f, lamC0 = 0.f, lamV0 = 0.f, dLamC = 0.f, dLamV = 0.f;
    double EC[NSUB], EV[NSUB];
#pragma unroll
    for (int s = 0; s < NSUB; s++)
        EC[s] = EV[s] = 0.0;
    long long Fx = 0, Fy = 0, Fz = 0;
    auto flushOctet = [&]() {
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
        if (ENERGY && iValid) {
#pragma unroll
            for (int u = 0; u < NSUB; u++) {
                sAccC[subI*NSUB+u][lane] += EC[u];
                sAccV[subI*NSUB+u][lane] += EV[u];
            }
        }
        Fx = Fy = Fz = 0;
#pragma unroll
        for (int s = 0; s < NSUB; s++)
            EC[s] = EV[s] = 0.0;
    };

    // Fix-up of one chunk: every lane moves its staged j atom (stage buffer `sb`, list entry in ring slot `ring`) into the
    // frame of the octet in flight and stores it into the pair planes at `pl`; tells whether the chunk is CLEAN -- no
    // exclusion bit, no padding entry (mask 0xff), every j atom in the same subset: such a chunk takes the lean copy of the
    // pair loop (no mask plane, one lambda pair and one energy pair for the whole chunk).
    // (in three parts, so that the early fix-up can be threaded through the pair loop: the shared-memory instructions are
    // volatile asm and keep their program order, the arithmetic between them is the compiler's to interleave)
    struct Staged { float4 pj, cJ, prm; unsigned mask; };
    auto fixupLoad = [&](unsigned sb, unsigned ring, Staged& t) {
        t.pj = St::ld4(sb+lane*St::STRIDE);
        t.cJ = ldSharedF4(sb+St::CTR+lane*16);
        t.prm = St::ld4(sb+St::PRM+lane*St::STRIDE);
        t.mask = ldSharedU32(entBase+ring*256u+4u);
    };
    auto fixupStore = [&](Staged& t, unsigned pl, bool perPairNow) {
        fixupShift<float>(t.pj, t.cJ, cI, oc, !perPairNow, wb);
        const float maskF = __uint_as_float(t.mask);
        stSharedF32(st1+pl+PL_XY, swapStores ? t.pj.y : t.pj.x); stSharedF32(st2+pl+PL_XY, swapStores ? t.pj.x : t.pj.y);
        stSharedF32(st1+pl+PL_ZQ, swapStores ? t.pj.w : t.pj.z); stSharedF32(st2+pl+PL_ZQ, swapStores ? t.pj.z : t.pj.w);
        stSharedF32(st1+pl+PL_SE, swapStores ? t.prm.y : t.prm.x); stSharedF32(st2+pl+PL_SE, swapStores ? t.prm.x : t.prm.y);
        stSharedF32(st1+pl+PL_SM, swapStores ? maskF : t.prm.z); stSharedF32(st2+pl+PL_SM, swapStores ? t.prm.z : maskF);
    };
    auto fixupVote = [&](const Staged& t, bool& cleanOut, float& subJ0Out) {
        subJ0Out = __shfl_sync(full, t.prm.z, 0);
        cleanOut = !__any_sync(full, t.mask != 0u || t.prm.z != subJ0Out);
    };
    auto fixup = [&](unsigned sb, unsigned ring, unsigned pl, bool perPairNow, bool& cleanOut, float& subJ0Out) {
        Staged t;
        fixupLoad(sb, ring, t);
        fixupStore(t, pl, perPairNow);
        fixupVote(t, cleanOut, subJ0Out);
    };

    // Pipeline (c = the chunk in the pair loop):  list entries of c+3 and atoms of c+2 in flight (cp.async), atoms of c+1
    // staged and FIXED UP DURING the pair loop of c -- the fix-up's chain of shared-memory loads, roundings and stores has
    // nothing to do with the pair arithmetic, so the two streams interleave (ncu on the version that fixed chunk c up at the
    // top of its own iteration: 29 % of the instructions but 41 % of the warp time in that stretch).  The early fix-up
    // uses the frame of the octet in flight; when chunk c+1 turns out to belong to the next octet it is simply redone at
    // the top of its iteration (one chunk in eleven).
    if (lane < 4)
        sOct[lane] = 0;
    __syncwarp();
    if (cBegin < cEnd) {
        fetchEntries(cBegin);
        if (cBegin+1 < cEnd)
            fetchEntries(cBegin+1);
        cpAsyncCommit();
        cpAsyncWait<0>();
        __syncwarp();
        gather(cBegin);
        if (cBegin+2 < cEnd)
            fetchEntries(cBegin+2);
        cpAsyncCommit();
        cpAsyncWait<0>();
        __syncwarp();
        if (cBegin+1 < cEnd)
            gather(cBegin+1);
        cpAsyncCommit();
    }
    bool cleanNext = false;
    float subJ0Next = 0.f;
    for (int ch = cBegin; ch < cEnd; ch++) {
        const unsigned sbNext = stageBase+(unsigned) ((ch+1)&1)*BUF;
        const unsigned pl = (unsigned) (ch&1)*PL_BYTES, plNext = (unsigned) ((ch+1)&1)*PL_BYTES;
        // the atoms of chunk ch+1 and the entries of chunk ch+2 were requested a whole chunk ago
        cpAsyncWait<0>();
        __syncwarp();                       // ... by every lane; and every lane is done with chunk ch-1
        const int octCur = __shfl_sync(full, (int) ldSharedU32(octBase+(unsigned) (ch&3)*4u), 0);
        // ---- octet switch
        const int O = octCur&0x3fffffff;
        bool clean = cleanNext;
        float subJ0 = subJ0Next;
        const bool perPair = (octCur&0x40000000) != 0;      // octet too wide for one image per j atom
        if (O != curO) {
            if (curO >= 0)
                flushOctet();
            curO = O;
            ki = O*8+k;
            pi = a.posqLocal[ki];
            si4 = a.sparams[ki];
            subI = (int) si4.z;
            iValid = subI >= 0;
            origI = a.sortedAtom[ki];
            cI = a.blockCenter[O>>2];
            oc = a.octetCenter[O];
            lamC0 = (float) a.sp->lambdaC[sliceOf(max(subI, 0), 0)];
            lamV0 = (float) a.sp->lambdaV[sliceOf(max(subI, 0), 0)];
            if (NSUB == 2) {
                dLamC = (float) a.sp->lambdaC[sliceOf(max(subI, 0), 1)]-lamC0;
                dLamV = (float) a.sp->lambdaV[sliceOf(max(subI, 0), 1)]-lamV0;
            }
            // a padding lane can never be "in range": its bounds are negative
            myCutHi = iValid ? (perPair ? c.cutoffHiWide : bandHi) : -1.f;
            myCutLo = iValid ? (perPair ? c.cutoffLoWide : bandLo) : -1.f;
            // the pair loop only has to NOTICE a band pair (bandRescan then classifies every pair of the chunk again with the
            // exact bounds): |r^2 - mid| < tol, a superset of [lo, hi) -- the difference is exact (Sterbenz), mid carries half
            // an ulp (3e-8 against a band of 2.6e-6 at least), the tolerance is widened by 5% + 1e-7 for it
            myCutMid = iValid ? 0.5f*(myCutHi+myCutLo) : -1.f;
            myBandTol = iValid ? fmaf(0.5f*(myCutHi-myCutLo), 1.05f, 1e-7f) : 0.f;
            // this chunk was fixed up in the previous octet's frame (or not at all): again, in its own (its atoms sit in the
            // stage buffer the NEXT gather is about to reuse, hence before that gather is issued)
            fixup(stageBase+(unsigned) (ch&1)*BUF, (unsigned) (ch&3), pl, perPair, clean, subJ0);
            __syncwarp();
        }
        if (ch+2 < cEnd) {
            gather(ch+2);
            if (ch+3 < cEnd)
                fetchEntries(ch+3);
        }
        cpAsyncCommit();
        float2 fix = bc(0.f), fiy = bc(0.f), fiz = bc(0.f), fjx = bc(0.f), fjy = bc(0.f), fjz = bc(0.f);
        float2 eCt = bc(0.f), eVt = bc(0.f), eC1 = bc(0.f), eV1 = bc(0.f);
        bool anyBand = false;
        // ---- 4 double-steps of 64 pairs ----
        auto steps = [&](auto perPairTag, auto cleanTag) {
        constexpr bool PERPAIR = decltype(perPairTag)::value;
        constexpr bool CLEAN = decltype(cleanTag)::value;
        // CLEAN: the lambdas of the chunk's one j subset
        const float lamCc = fmaf(subJ0, dLamC, lamC0), lamVc = fmaf(subJ0, dLamV, lamV0);
        // the next chunk's fix-up, threaded through the pair arithmetic (see above): its loads first, its stores after the
        // second double-step's loads
        Staged nextAtoms;
        fixupLoad(sbNext, (unsigned) ((ch+1)&3), nextAtoms);
#pragma unroll
        for (int d = 0; d < 4; d++) {
            if (d == 2)
                fixupStore(nextAtoms, plNext, PERPAIR);
            const unsigned ad = (ownPair+pl)^(unsigned) (d*16);
            const float4 xy = ldSharedF4(ad+PL_XY), zq = ldSharedF4(ad+PL_ZQ);
            const float4 se = ldSharedF4(ad+PL_SE);
            float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!CLEAN)
                sm = ldSharedF4(ad+PL_SM);
            const float2 m1 = bc(-1.f);
            float2 dx = fma2(make_float2(xy.x, xy.y), m1, bc(pi.x));
            float2 dy = fma2(make_float2(xy.z, xy.w), m1, bc(pi.y));
            float2 dz = fma2(make_float2(zq.x, zq.y), m1, bc(pi.z));
            if (PERPAIR) {
                float2 t = mul2(dx, bc(wb.iax));
                dx = fma2(bc(-wb.ax), make_float2(rintf(t.x), rintf(t.y)), dx);
                t = mul2(dy, bc(wb.iby));
                dy = fma2(bc(-wb.by), make_float2(rintf(t.x), rintf(t.y)), dy);
                t = mul2(dz, bc(wb.icz));
                dz = fma2(bc(-wb.cz), make_float2(rintf(t.x), rintf(t.y)), dz);
            }
            const float2 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            const bool ok0 = CLEAN || !(__float_as_uint(sm.z)&kBit), ok1 = CLEAN || !(__float_as_uint(sm.w)&kBit);
            const bool in0 = ok0 && r2.x < myCutLo, in1 = ok1 && r2.y < myCutLo;
            const float2 off = add2(r2, bc(-myCutMid));
            anyBand |= (ok0 & (fabsf(off.x) < myBandTol)) | (ok1 & (fabsf(off.y) < myBandTol));
            // a masked pair computes on r^2 = 1e10 nm^2: exp(-a^2 r^2) = 0 kills the Coulomb term exactly, the
            // Lennard-Jones force is ~1e-43 of a real one
            const float2 r2m = make_float2(in0 ? r2.x : 1e10f, in1 ? r2.y : 1e10f);
            // 1/r: MUFU.RSQ (2 ulp) + one Newton step (the LJ force goes as r^-14)
            const float2 y = make_float2(fastRsqrt(r2m.x), fastRsqrt(r2m.y));
            const float2 invR = mul2(y, fma2(mul2(r2m, bc(-0.5f)), mul2(y, y), bc(1.5f)));
            const float2 invR2 = mul2(invR, invR);
            const float2 r = mul2(r2m, invR);
            // Coulomb: qq/r exp(-a^2 r^2) (erfcx(ar), erfcx(ar) + 2 a r/sqrt(pi))
            const float2 e = mul2(r2m, bc(c.negAlpha2Log2e));
            const float2 ex = make_float2(fastExp2(e.x), fastExp2(e.y));
            const float2 w = fma2(r, bc(c.halfAlpha), bc(1.f));
            const float2 u = fma2(make_float2(fastRcp(w.x), fastRcp(w.y)), bc(3.f), bc(-2.f));
#ifdef PACKED_ESTRIN
            // tuning variant (not the default; unmeasured): Estrin's scheme, 12 operations in a chain of 4 instead of 9 in a chain of 9
            const float2 u2 = mul2(u, u), u4 = mul2(u2, u2), u8 = mul2(u4, u4);
            const float2 q01 = fma2(bc(4.0981802125e-01f), u, bc(4.2758357745e-01f)), q23 = fma2(bc(2.2509531568e-02f), u, bc(1.4242693719e-01f));
            const float2 q45 = fma2(bc(-8.6138957608e-04f), u, bc(-1.6015017984e-03f)), q67 = fma2(bc(3.4861879002e-05f), u, bc(9.8919924926e-05f));
            const float2 q89 = fma2(bc(-7.5385177448e-07f), u, bc(-8.2052179778e-06f));
            const float2 p = fma2(q89, u8, fma2(fma2(q67, u2, q45), u4, fma2(q23, u2, q01)));
#else
            float2 p = fma2(bc(-7.5385177448e-07f), u, bc(-8.2052179778e-06f));
            p = fma2(p, u, bc(3.4861879002e-05f));
            p = fma2(p, u, bc(9.8919924926e-05f));
            p = fma2(p, u, bc(-8.6138957608e-04f));
            p = fma2(p, u, bc(-1.6015017984e-03f));
            p = fma2(p, u, bc(2.2509531568e-02f));
            p = fma2(p, u, bc(1.4242693719e-01f));
            p = fma2(p, u, bc(4.0981802125e-01f));
            p = fma2(p, u, bc(4.2758357745e-01f));
#endif
            const float2 g = mul2(mul2(mul2(make_float2(zq.z, zq.w), bc(pi.w)), invR), ex);
            const float2 ec = mul2(g, p);
            const float2 fc = mul2(g, fma2(bc(c.twoAlphaOverSqrtPi), r, p));
            // Lennard-Jones
            float2 s2 = mul2(add2(make_float2(se.x, se.y), bc(si4.x)), invR);
            s2 = mul2(s2, s2);
            const float2 s6 = mul2(mul2(s2, s2), s2);
            const float2 es6 = mul2(mul2(make_float2(se.z, se.w), bc(si4.y)), s6);
            const float2 fv = mul2(es6, fma2(bc(12.f), s6, bc(-6.f)));
            const float2 ev = fma2(es6, s6, neg2(es6));
            // lambda-weighted force over r
            float2 f;
            const float2 subJ = make_float2(sm.x, sm.y);
            if (NSUB == 2 && !CLEAN)
                f = mul2(fma2(fma2(subJ, bc(dLamV), bc(lamV0)), fv, mul2(fma2(subJ, bc(dLamC), bc(lamC0)), fc)), invR2);
            else if (NSUB == 2)
                f = mul2(fma2(bc(lamVc), fv, mul2(bc(lamCc), fc)), invR2);
            else
                f = mul2(fma2(bc(lamV0), fv, mul2(bc(lamC0), fc)), invR2);
            fix = fma2(f, dx, fix); fiy = fma2(f, dy, fiy); fiz = fma2(f, dz, fiz);
            fjx = fma2(f, dx, fjx); fjy = fma2(f, dy, fjy); fjz = fma2(f, dz, fjz);     // sign applied once, at the end
            if (ENERGY) {
                eCt = add2(eCt, ec); eVt = add2(eVt, ev);
                if (NSUB == 2 && !CLEAN) {
                    eC1 = fma2(subJ, ec, eC1); eV1 = fma2(subJ, ev, eV1);
                }
            }
            if (DUMP && (in0 || in1)) {
                const unsigned slot0 = (((ad-pairBase-pl)>>4)<<1);
#pragma unroll
                for (int h = 0; h < 2; h++)
                    if (h == 0 ? in0 : in1) {
                        const int origJ = a.sortedAtom[sEnt[ch&3][slot0+h].x];
                        unsigned long long idx = atomicAdd(a.pairDumpCount, 1ull);
                        if ((long long) idx < a.pairDumpCapacity)
                            a.pairDump[idx] = make_int2(min(origI, origJ), max(origI, origJ));
                    }
            }
            // the accumulators follow their slot pair to the lane that takes it next (home after the last)
            const int hop = (d&1) ? 6 : 2;
            fjx.x = __shfl_xor_sync(full, fjx.x, hop); fjx.y = __shfl_xor_sync(full, fjx.y, hop);
            fjy.x = __shfl_xor_sync(full, fjy.x, hop); fjy.y = __shfl_xor_sync(full, fjy.y, hop);
            fjz.x = __shfl_xor_sync(full, fjz.x, hop); fjz.y = __shfl_xor_sync(full, fjz.y, hop);
        }
        fixupVote(nextAtoms, cleanNext, subJ0Next);
        };
        if (perPair)
            steps(std::true_type(), std::false_type());     // rare (octets wider than half the box minus the cutoff): one copy
        else if (clean)
            steps(std::false_type(), std::true_type());
        else
            steps(std::false_type(), std::false_type());
        if (ENERGY && NSUB == 2 && clean && !perPair) {
            // the chunk's energies belong to its one j subset
            eC1 = mul2(eCt, bc(subJ0)); eV1 = mul2(eVt, bc(subJ0));
        }
        if (__any_sync(full, anyBand)) {
            if (anyBand) {
                DirectOut o;
                o.posd = a.posd; o.sortedAtom = a.sortedAtom; o.sp = a.sp; o.forceSorted = a.forceSorted; o.paddedAtoms = a.paddedAtoms;
                o.energyTable = a.energyBuf+(size_t) (blockIdx.x%ENERGY_REPLICAS)*2*SNB_MAX_SLICES;
                o.pairDump = a.pairDump; o.pairDumpCount = a.pairDumpCount; o.pairDumpCapacity = a.pairDumpCapacity;
                bandRescanPlanes<ENERGY, DUMP>(o, c, lane, ki, origI, myCutLo, myCutHi, pi, si4, perPair ? 1 : 0, &sPairs[ch&1][0], &sEnt[ch&3][0]);
            }
        }
        // lanes k and k^1 hold the two partial sums of the same slot pair: the even lane keeps slot 2P, the odd one 2P+1
        {
            const float sx = odd ? fjx.x : fjx.y, sy = odd ? fjy.x : fjy.y, sz = odd ? fjz.x : fjz.y;
            const float jx = (odd ? fjx.y : fjx.x)+__shfl_xor_sync(full, sx, 1);
            const float jy = (odd ? fjy.y : fjy.x)+__shfl_xor_sync(full, sy, 1);
            const float jz = (odd ? fjz.y : fjz.x)+__shfl_xor_sync(full, sz, 1);
            const int jCur = (int) ldSharedU32(entBase+(unsigned) (ch&3)*256u);
            if (jCur >= 0) {
                addFixed(&a.forceSorted[jCur], toFixed(-jx));
                addFixed(&a.forceSorted[a.paddedAtoms+jCur], toFixed(-jy));
                addFixed(&a.forceSorted[2*a.paddedAtoms+jCur], toFixed(-jz));
            }
        }
        Fx += toFixed(fix.x+fix.y); Fy += toFixed(fiy.x+fiy.y); Fz += toFixed(fiz.x+fiz.y);
        if (ENERGY) {
            const float c1 = eC1.x+eC1.y, v1 = eV1.x+eV1.y;
            if (NSUB == 2) {
                EC[0] += (double) ((eCt.x+eCt.y)-c1); EC[NSUB-1] += (double) c1;
                EV[0] += (double) ((eVt.x+eVt.y)-v1); EV[NSUB-1] += (double) v1;
            }
            else {
                EC[0] += (double) (eCt.x+eCt.y);
                EV[0] += (double) (eVt.x+eVt.y);
            }
        }
    }
    if (curO >= 0)
        flushOctet();
    if (ENERGY) {
        double* table = a.energyBuf+(size_t) (blockIdx.x%ENERGY_REPLICAS)*2*SNB_MAX_SLICES;
#pragma unroll
        for (int s = 0; s < NSUB; s++)
#pragma unroll
            for (int u = 0; u < NSUB; u++) {
                double ecs = sAccC[s*NSUB+u][lane], evs = sAccV[s*NSUB+u][lane];
                for (int o = 16; o > 0; o >>= 1) {
                    ecs += __shfl_xor_sync(full, ecs, o);
                    evs += __shfl_xor_sync(full, evs, o);
                }
                if (lane == 0) {
                    if (ecs != 0.0) atomicAdd(&table[2*sliceOf(s, u)], ecs);
                    if (evs != 0.0) atomicAdd(&table[2*sliceOf(s, u)+1], evs);
                }
            }
    }
}

template <int NSUB> void launchPackedN(const DirectArgs& a, cudaStream_t st, int grid) {
    if (a.pairDump)
        k_directPacked<NSUB, true, true><<<grid, 32, 0, st>>>(a);
    else if (a.needEnergy)
        k_directPacked<NSUB, true, false><<<grid, 32, 0, st>>>(a);
    else
        k_directPacked<NSUB, false, false><<<grid, 32, 0, st>>>(a);
}

}  // namespace

// Takes the launch if this is the packed kernel's case; false = the caller launches k_direct.
bool launchDirectPacked(const DirectArgs& a, cudaStream_t st, int grid) {
    if (a.doublePrecision || a.kind != KIND_EWALD || a.useSwitch || a.numSubsets > 2)
        return false;
    if (a.p.pbcMode != PBC_SHIFT && a.p.pbcMode != PBC_PAIR_ORTHO)
        return false;
    grid = grid/DIRECT_CTAS_PER_SM*PACKED_CTAS_PER_SM;
    if (a.numSubsets == 2)
        launchPackedN<2>(a, st, grid);
    else
        launchPackedN<1>(a, st, grid);
    return true;
}
