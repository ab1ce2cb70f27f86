// Tile-list build (north-star subsystem 1b).
//
// The sorted atoms are cut into blocks of 32 (j side) and octets of 8 (i side; 4 per block).  For
// every octet the list holds the later atoms that come within the cutoff of the octet's bounding
// box -- the tile interaction flags, compacted into chunks of 32 j atoms -- each with an 8-bit
// exclusion mask over the octet's atoms.  An octet's box is much tighter than a 32-atom block's, so
// ~40% of the pairs a tile visits are in range instead of ~27%.
//
//  k_tileFlags   one warp per (i block, group of 32 candidate j blocks): block-box test per lane, then
//                for every hit block four 32-bit words of per-atom flags, one per octet of the i block.
//  k_compactList one warp per octet: prefix sums of the popcounts place every flagged j atom in the
//                octet's contiguous chunk range; exclusion masks are OR-ed in from the octet's own
//                (short) exclusion lists.
// No order-dependent atomics: chunk CONTENTS are deterministic, only their location in the list varies.
//
// No counterpart source exists in the reference tree (OpenMM's CudaNonbondedUtilities does this job
// there, un-vendored); exclusion semantics: every exception excludes its pair
// (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:104-112).
#include "kernels.cuh"

#define FLAG_WARPS 8
#define COMPACT_WARPS 4

__device__ __forceinline__ float wrapOrtho(float d, float L, float invL) { return d-L*rintf(d*invL); }

__device__ __forceinline__ float boxDist2(float dx, float dy, float dz, float3 h) {
    float x = fmaxf(0.f, fabsf(dx)-h.x), y = fmaxf(0.f, fabsf(dy)-h.y), z = fmaxf(0.f, fabsf(dz)-h.z);
    return x*x+y*y+z*z;
}

// Smallest box distance over the periodic images the pair kernel can pick.
__device__ __forceinline__ float imageDist2(float dx, float dy, float dz, float3 h, const ListArgs& a) {
    if (a.pbcMode == PBC_NONE)
        return boxDist2(dx, dy, dz, h);
    if (a.pbcMode != PBC_PAIR_TRICLINIC) {
        dx = wrapOrtho(dx, a.box.ax, a.box.invAx);
        dy = wrapOrtho(dy, a.box.by, a.box.invBy);
        dz = wrapOrtho(dz, a.box.cz, a.box.invCz);
        return boxDist2(dx, dy, dz, h);
    }
    // triclinic: reduce like the pair kernel does, then try the 27 neighbouring images
    float s = rintf(dz*a.box.invCz);
    dx -= s*a.box.cx; dy -= s*a.box.cy; dz -= s*a.box.cz;
    s = rintf(dy*a.box.invBy);
    dx -= s*a.box.bx; dy -= s*a.box.by;
    s = rintf(dx*a.box.invAx);
    dx -= s*a.box.ax;
    float best = 3.0e38f;
    for (int k = -1; k <= 1; k++)
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                float x = dx+i*a.box.ax+j*a.box.bx+k*a.box.cx;
                float y = dy+j*a.box.by+k*a.box.cy;
                float z = dz+k*a.box.cz;
                best = fminf(best, boxDist2(x, y, z, h));
            }
    return best;
}

// candidate c of i block I: the block itself and all later blocks
__device__ __forceinline__ int candidateBlock(const ListArgs& a, int I, int c) {
    const int J = I+c;
    return J < a.numBlocks ? J : -1;
}
__device__ __forceinline__ int candidateIndex(const ListArgs& a, int I, int J) { return J-I; }

__global__ void __launch_bounds__(FLAG_WARPS*32) k_tileFlags(ListArgs a) {
    const int lane = threadIdx.x&31;
    const int groups = a.flagStride>>5;
    const long long warpId = ((long long) blockIdx.x*blockDim.x+threadIdx.x)>>5;
    if (warpId >= (long long) a.numBlocks*groups)
        return;
    const int I = (int) (warpId/groups), g = (int) (warpId-(long long) I*groups);
    if (candidateBlock(a, I, g<<5) < 0)
        return;
    const unsigned full = 0xffffffffu;
    const float cut2 = a.cutoff*a.cutoff;
    const double cIx = a.blockCenter[3*I], cIy = a.blockCenter[3*I+1], cIz = a.blockCenter[3*I+2];
    const float4 e4 = a.blockExt[I];
    const float3 hI = make_float3(e4.x, e4.y, e4.z);
    float4 oc[4], oe[4];
#pragma unroll
    for (int o = 0; o < 4; o++) {
        oc[o] = a.octetCenter[4*I+o];
        oe[o] = a.octetExt[4*I+o];
    }
    const int c = (g<<5)+lane;
    const int J = candidateBlock(a, I, c);
    bool hit = J >= 0;
    if (hit && a.useCutoff && J != I) {
        const float4 eJ = a.blockExt[J];
        const float dx = (float) (a.blockCenter[3*J]-cIx), dy = (float) (a.blockCenter[3*J+1]-cIy), dz = (float) (a.blockCenter[3*J+2]-cIz);
        hit = imageDist2(dx, dy, dz, make_float3(hI.x+eJ.x, hI.y+eJ.y, hI.z+eJ.z), a) < cut2;
    }
    unsigned bits = __ballot_sync(full, hit);
    unsigned myWord[4] = {0u, 0u, 0u, 0u};
    while (bits) {
        const int b = __ffs(bits)-1;
        bits &= bits-1;
        const int Jb = __shfl_sync(full, J, b), j = Jb*32+lane;
        const bool valid = a.sortedAtom[j] >= 0;
        const float4 pj = a.posqLocal[j];
        const float px = pj.x+(float) (a.blockCenter[3*Jb]-cIx), py = pj.y+(float) (a.blockCenter[3*Jb+1]-cIy), pz = pj.z+(float) (a.blockCenter[3*Jb+2]-cIz);
#pragma unroll
        for (int o = 0; o < 4; o++) {
            // only atoms after the octet's own (every pair is visited once); the octet's own atoms are
            // the first entries of its list (k_compactList)
            bool acc = valid && oe[o].x >= 0.f && j > (4*I+o)*8+7;
            if (acc && a.useCutoff)
                acc = imageDist2(px-oc[o].x, py-oc[o].y, pz-oc[o].z, make_float3(oe[o].x, oe[o].y, oe[o].z), a) < cut2;
            const unsigned word = __ballot_sync(full, acc);
            if (lane == b)
                myWord[o] = word;
        }
    }
    if (c < a.flagStride)
#pragma unroll
        for (int o = 0; o < 4; o++)
            a.tileFlags[(size_t) (4*I+o)*a.flagStride+c] = myWord[o];
}

__global__ void __launch_bounds__(COMPACT_WARPS*32) k_compactList(ListArgs a) {
    extern __shared__ int sBase[];                  // [COMPACT_WARPS][flagStride] list-relative start of every candidate
    const int lane = threadIdx.x&31, w = threadIdx.x>>5;
    const int O = blockIdx.x*COMPACT_WARPS+w;       // octet
    if (O >= 4*a.numBlocks)
        return;
    const float4 ext = a.octetExt[O];
    if (ext.x < 0.f)
        return;                                     // padding octet
    const int I = O>>2;
    int* base = sBase+(size_t) w*a.flagStride;
    const unsigned full = 0xffffffffu;
    const unsigned* row = a.tileFlags+(size_t) O*a.flagStride;
    const int numCand = a.numBlocks-I;
    // pass 1: positions (the octet's own atoms occupy entries 0..7)
    int running = 8;
    for (int c0 = 0; c0 < numCand; c0 += 32) {
        const int c = c0+lane;
        const unsigned word = c < numCand ? row[c] : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        if (c < numCand)
            base[c] = running+incl-cnt;
        running += __shfl_sync(full, incl, 31);
    }
    const int total = running, n = (total+31)>>5;
    const int numItemsHere = (n+a.itemChunks-1)/a.itemChunks;
    int c0g = 0, item0 = 0;
    if (lane == 0) {
        c0g = atomicAdd(&a.counters[0], n);
        item0 = atomicAdd(&a.counters[1], numItemsHere);
        if (c0g+n > a.maxChunks || item0+numItemsHere > a.maxItems)
            a.counters[2] = 1;                      // overflow: the host grows the buffers and repeats
    }
    c0g = __shfl_sync(full, c0g, 0);
    item0 = __shfl_sync(full, item0, 0);
    if (c0g+n > a.maxChunks || item0+numItemsHere > a.maxItems)
        return;
    int2* out = a.jList+(size_t) c0g*32;
    // octets too wide for the one-image-per-j-atom shortcut wrap every pair instead
    const int wide = (a.pbcMode == PBC_SHIFT && (ext.x >= a.shiftLimit[0] || ext.y >= a.shiftLimit[1] || ext.z >= a.shiftLimit[2])) ? 1 : 0;
    for (int k = lane; k < numItemsHere; k += 32)
        a.workItems[item0+k] = make_int4(O, c0g+k*a.itemChunks, min(a.itemChunks, n-k*a.itemChunks), wide);
    // pass 2: entries.  Own atoms first (lower triangle masked: every pair once), then the tail padding
    const int ki = O*8+(lane&7);
    const int atomI = lane < 8 ? a.sortedAtom[ki] : -1;
    if (lane < 8)
        out[lane] = atomI >= 0 ? make_int2(ki, (int) ((0xffu<<lane)&0xffu)) : make_int2(-1, 0xff);
    for (int k = total+lane; k < n*32; k += 32)
        out[k] = make_int2(-1, 0xff);
    for (int c0 = 0; c0 < numCand; c0 += 32) {
        const int c = c0+lane;
        const unsigned word = c < numCand ? row[c] : 0u;
        unsigned hits = __ballot_sync(full, word != 0u);
        while (hits) {
            const int b = __ffs(hits)-1;
            hits &= hits-1;
            const unsigned wb = __shfl_sync(full, word, b);
            const int J = candidateBlock(a, I, c0+b);
            if ((wb>>lane)&1u)
                out[base[c0+b]+__popc(wb&((1u<<lane)-1))] = make_int2(J*32+lane, 0);
        }
    }
    __syncwarp();
    // pass 3: exclusion masks from the i side; partners before the octet are that octet's business
    if (atomI >= 0) {
        const int e0 = a.exclStart[atomI], e1 = a.exclStart[atomI+1];
        for (int e = e0; e < e1; e++) {
            const int sp = a.sortedPos[a.exclList[e]];
            if ((sp>>3) == O) {
                if ((sp&7) > lane)
                    atomicOr(&out[sp&7].y, (int) (1u<<lane));
            }
            else if (sp > O*8+7) {
                const int c = candidateIndex(a, I, sp>>5), b = sp&31;
                const unsigned word = row[c];
                if ((word>>b)&1u)
                    atomicOr(&out[base[c]+__popc(word&((1u<<b)-1))].y, (int) (1u<<lane));
            }
        }
    }
}

void launchBuildList(const ListArgs& a, cudaStream_t stream, int numSMs) {
    (void) numSMs;
    const int groups = a.flagStride>>5;
    const long long warps = (long long) a.numBlocks*groups;
    k_tileFlags<<<(unsigned) ((warps+FLAG_WARPS-1)/FLAG_WARPS), FLAG_WARPS*32, 0, stream>>>(a);
    const size_t smem = sizeof(int)*(size_t) COMPACT_WARPS*a.flagStride;
    static bool attrSet = false;
    if (!attrSet) {
        cudaFuncSetAttribute(k_compactList, cudaFuncAttributeMaxDynamicSharedMemorySize, 4*COMPACT_WARPS*SNB_FLAGS_MAX_BLOCKS);
        attrSet = true;
    }
    k_compactList<<<(4*a.numBlocks+COMPACT_WARPS-1)/COMPACT_WARPS, COMPACT_WARPS*32, smem, stream>>>(a);
}
