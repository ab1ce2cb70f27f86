// Tile-list build (north-star subsystem 1b): for every block of 32 sorted atoms (the "i block")
// find the later blocks whose bounding boxes come within the cutoff, flag which of their atoms
// come within the cutoff of the i block's box (the tile interaction flags), compact the flagged
// atoms into chunks of 32, and attach to every j atom a 32-bit exclusion mask over the i lanes.
//
// No counterpart source exists in the reference tree (OpenMM's CudaNonbondedUtilities does this
// job there, un-vendored); the exclusion semantics follow
// platforms/reference/src/ReferenceOpenMMLabKernels.cpp:104-112 (every exception excludes its pair).
#include "kernels.cuh"

#define LIST_WARPS 4
#define BUF_ENTRIES (SNB_ITEM_CHUNKS*32+32)
#define FLAG_WARPS 8
#define COMPACT_WARPS 4

__device__ __forceinline__ float wrapOrtho(float d, float L, float invL) { return d-L*rintf(d*invL); }

// squared distance between a point/box at offset d (half extents hJ) and the i box (half extents hI)
__device__ __forceinline__ float boxDist2(float dx, float dy, float dz, float3 h) {
    float x = fmaxf(0.f, fabsf(dx)-h.x), y = fmaxf(0.f, fabsf(dy)-h.y), z = fmaxf(0.f, fabsf(dz)-h.z);
    return x*x+y*y+z*z;
}

// Smallest box distance over the periodic images allowed by the mode.
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

__device__ __forceinline__ unsigned exclusionMask(int atomJ, int blockI, const ListArgs& a) {
    unsigned m = 0;
    int e0 = a.exclStart[atomJ], e1 = a.exclStart[atomJ+1];
    for (int e = e0; e < e1; e++) {
        int sp = a.sortedPos[a.exclList[e]];
        if ((sp>>5) == blockI)
            m |= 1u<<(sp&31);
    }
    return m;
}

// Write `n` full chunks from the warp's buffer to the global list as one work item.
__device__ __forceinline__ void flushChunks(const ListArgs& a, int blockI, const int2* buf, int n, int lane, int wide) {
    int base = 0, item = 0;
    if (lane == 0) {
        base = atomicAdd(&a.counters[0], n);
        item = atomicAdd(&a.counters[1], 1);
        if (base+n > a.maxChunks || item >= a.maxItems)
            a.counters[2] = 1;          // overflow: the host grows the buffers and repeats the evaluation
        else
            a.workItems[item] = make_int4(blockI, base, n, wide);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (base+n > a.maxChunks || item >= a.maxItems)
        return;
    for (int c = 0; c < n; c++)
        a.jList[(size_t) (base+c)*32+lane] = buf[c*32+lane];
}

__global__ void __launch_bounds__(LIST_WARPS*32) k_buildList(ListArgs a) {
    __shared__ int2 sbuf[LIST_WARPS][BUF_ENTRIES];
    int lane = threadIdx.x&31, w = threadIdx.x>>5;
    int2* buf = sbuf[w];
    const unsigned full = 0xffffffffu;
    float cut2 = a.cutoff*a.cutoff;
    for (int I = blockIdx.x*LIST_WARPS+w; I < a.numBlocks; I += gridDim.x*LIST_WARPS) {
        double cIx = a.blockCenter[3*I], cIy = a.blockCenter[3*I+1], cIz = a.blockCenter[3*I+2];
        float4 e4 = a.blockExt[I];
        float3 hI = make_float3(e4.x, e4.y, e4.z);
        const int wide = (a.pbcMode == PBC_SHIFT && (e4.x >= a.shiftLimit[0] || e4.y >= a.shiftLimit[1] || e4.z >= a.shiftLimit[2])) ? 1 : 0;
        // diagonal chunk: j = own atoms, lower triangle masked so that every pair is visited once
        int j = I*32+lane;
        int atomJ = a.sortedAtom[j];
        unsigned mask = full<<lane;
        if (atomJ >= 0)
            mask |= exclusionMask(atomJ, I, a);
        buf[lane] = atomJ >= 0 ? make_int2(j, (int) mask) : make_int2(-1, -1);
        int count = 32;
        for (int J0 = I+1; J0 < a.numBlocks; J0 += 32) {
            int J = J0+lane;
            bool hit = false;
            if (J < a.numBlocks) {
                hit = true;
                if (a.useCutoff) {
                    float4 eJ = a.blockExt[J];
                    float dx = (float) (a.blockCenter[3*J]-cIx), dy = (float) (a.blockCenter[3*J+1]-cIy), dz = (float) (a.blockCenter[3*J+2]-cIz);
                    hit = imageDist2(dx, dy, dz, make_float3(hI.x+eJ.x, hI.y+eJ.y, hI.z+eJ.z), a) < cut2;
                }
            }
            unsigned bits = __ballot_sync(full, hit);
            while (bits) {
                int b = __ffs(bits)-1;
                bits &= bits-1;
                int Jb = J0+b;
                j = Jb*32+lane;
                atomJ = a.sortedAtom[j];
                bool acc = atomJ >= 0;
                if (acc && a.useCutoff) {
                    float4 pj = a.posqLocal[j];
                    float dx = pj.x+(float) (a.blockCenter[3*Jb]-cIx), dy = pj.y+(float) (a.blockCenter[3*Jb+1]-cIy), dz = pj.z+(float) (a.blockCenter[3*Jb+2]-cIz);
                    acc = imageDist2(dx, dy, dz, hI, a) < cut2;
                }
                unsigned m = __ballot_sync(full, acc);
                if (acc) {
                    unsigned ex = exclusionMask(atomJ, I, a);
                    buf[count+__popc(m&((1u<<lane)-1))] = make_int2(j, (int) ex);
                }
                count += __popc(m);
                __syncwarp();
                if (count >= a.itemChunks*32) {
                    flushChunks(a, I, buf, a.itemChunks, lane, wide);
                    __syncwarp();
                    int2 rest = buf[a.itemChunks*32+lane];
                    __syncwarp();
                    buf[lane] = rest;
                    count -= a.itemChunks*32;
                    __syncwarp();
                }
            }
        }
        // tail: pad the last chunk with empty entries
        int n = (count+31)/32;
        for (int k = count+lane; k < n*32; k += 32)
            buf[k] = make_int2(-1, -1);
        __syncwarp();
        if (n > 0)
            flushChunks(a, I, buf, n, lane, wide);
        __syncwarp();
    }
}

// =====================================================================================
// Parallel two-kernel build for small and medium systems (numBlocks <= SNB_FLAGS_MAX_BLOCKS),
// where one warp per i block would leave the GPU mostly idle (DHFR: 736 blocks on 148 SMs).
//
//  k_tileFlags   one warp per (i block, group of 32 later blocks): bounding-box test per lane,
//                then for every hit block a 32-bit word of per-atom interaction flags.
//  k_compactList one warp per i block: prefix sums of the flag words' popcounts place every flagged
//                j atom in the block's contiguous chunk range; the exclusion masks are then OR-ed in
//                from the i atoms' own (short) exclusion lists.
// Both are free of order-dependent atomics: chunk CONTENTS are deterministic, only their location
// in the list varies from run to run.
// =====================================================================================
__global__ void __launch_bounds__(FLAG_WARPS*32) k_tileFlags(ListArgs a) {
    const int lane = threadIdx.x&31;
    const int groups = (a.numBlocks+31)>>5;
    const long long warpId = ((long long) blockIdx.x*blockDim.x+threadIdx.x)>>5;
    if (warpId >= (long long) a.numBlocks*groups)
        return;
    const int I = (int) (warpId/groups), g = (int) (warpId-(long long) I*groups);
    const int J0 = g<<5;
    if (J0+31 <= I)
        return;                                     // nothing above the diagonal in this group
    const unsigned full = 0xffffffffu;
    const float cut2 = a.cutoff*a.cutoff;
    const double cIx = a.blockCenter[3*I], cIy = a.blockCenter[3*I+1], cIz = a.blockCenter[3*I+2];
    const float4 e4 = a.blockExt[I];
    const float3 hI = make_float3(e4.x, e4.y, e4.z);
    const int J = J0+lane;
    bool hit = J > I && J < a.numBlocks;
    if (hit && a.useCutoff) {
        const float4 eJ = a.blockExt[J];
        const float dx = (float) (a.blockCenter[3*J]-cIx), dy = (float) (a.blockCenter[3*J+1]-cIy), dz = (float) (a.blockCenter[3*J+2]-cIz);
        hit = imageDist2(dx, dy, dz, make_float3(hI.x+eJ.x, hI.y+eJ.y, hI.z+eJ.z), a) < cut2;
    }
    unsigned bits = __ballot_sync(full, hit);
    unsigned myWord = 0;
    while (bits) {
        const int b = __ffs(bits)-1;
        bits &= bits-1;
        const int Jb = J0+b, j = Jb*32+lane;
        bool acc = a.sortedAtom[j] >= 0;
        if (acc && a.useCutoff) {
            const float4 pj = a.posqLocal[j];
            const float dx = pj.x+(float) (a.blockCenter[3*Jb]-cIx), dy = pj.y+(float) (a.blockCenter[3*Jb+1]-cIy), dz = pj.z+(float) (a.blockCenter[3*Jb+2]-cIz);
            acc = imageDist2(dx, dy, dz, hI, a) < cut2;
        }
        const unsigned word = __ballot_sync(full, acc);
        if (lane == b)
            myWord = word;
    }
    if (J < a.numBlocks)
        a.tileFlags[(size_t) I*a.flagStride+J] = myWord;
}

__global__ void __launch_bounds__(COMPACT_WARPS*32) k_compactList(ListArgs a) {
    extern __shared__ int sBase[];                  // [COMPACT_WARPS][flagStride] chunk-relative start of every j block
    const int lane = threadIdx.x&31, w = threadIdx.x>>5;
    const int I = blockIdx.x*COMPACT_WARPS+w;
    if (I >= a.numBlocks)
        return;
    int* base = sBase+(size_t) w*a.flagStride;
    const unsigned full = 0xffffffffu;
    const unsigned* row = a.tileFlags+(size_t) I*a.flagStride;
    // pass 1: positions (the diagonal chunk occupies entries 0..31)
    int running = 32;
    for (int J0 = (I+1)&~31; J0 < a.numBlocks; J0 += 32) {
        const int J = J0+lane;
        const unsigned word = (J > I && J < a.numBlocks) ? row[J] : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        if (J < a.numBlocks)
            base[J] = running+incl-cnt;
        running += __shfl_sync(full, incl, 31);
    }
    const int total = running, n = (total+31)>>5;
    const int numItemsHere = (n+a.itemChunks-1)/a.itemChunks;
    int c0 = 0, item0 = 0;
    if (lane == 0) {
        c0 = atomicAdd(&a.counters[0], n);
        item0 = atomicAdd(&a.counters[1], numItemsHere);
        if (c0+n > a.maxChunks || item0+numItemsHere > a.maxItems)
            a.counters[2] = 1;
    }
    c0 = __shfl_sync(full, c0, 0);
    item0 = __shfl_sync(full, item0, 0);
    if (c0+n > a.maxChunks || item0+numItemsHere > a.maxItems)
        return;
    int2* out = a.jList+(size_t) c0*32;
    // per-item flag: this i block is too wide for the one-image-per-j-atom shortcut
    const float4 e4 = a.blockExt[I];
    const int wide = (a.pbcMode == PBC_SHIFT && (e4.x >= a.shiftLimit[0] || e4.y >= a.shiftLimit[1] || e4.z >= a.shiftLimit[2])) ? 1 : 0;
    for (int k = lane; k < numItemsHere; k += 32)
        a.workItems[item0+k] = make_int4(I, c0+k*a.itemChunks, min(a.itemChunks, n-k*a.itemChunks), wide);
    // pass 2: entries.  Diagonal chunk first (lower triangle masked so that every pair is visited once)
    const int ki = I*32+lane;
    const int atomI = a.sortedAtom[ki];
    out[lane] = atomI >= 0 ? make_int2(ki, (int) (full<<lane)) : make_int2(-1, -1);
    for (int k = total+lane; k < n*32; k += 32)
        out[k] = make_int2(-1, -1);
    for (int J0 = (I+1)&~31; J0 < a.numBlocks; J0 += 32) {
        const int J = J0+lane;
        const unsigned word = (J > I && J < a.numBlocks) ? row[J] : 0u;
        unsigned hits = __ballot_sync(full, word != 0u);
        while (hits) {
            const int b = __ffs(hits)-1;
            hits &= hits-1;
            const unsigned wb = __shfl_sync(full, word, b);
            const int start = base[J0+b];
            if ((wb>>lane)&1u)
                out[start+__popc(wb&((1u<<lane)-1))] = make_int2((J0+b)*32+lane, 0);
        }
    }
    __syncwarp();
    // pass 3: exclusion masks from the i side (every exception excludes its pair,
    // ReferenceOpenMMLabKernels.cpp:104-112); partners in earlier blocks are that block's business
    if (atomI >= 0) {
        const int e0 = a.exclStart[atomI], e1 = a.exclStart[atomI+1];
        for (int e = e0; e < e1; e++) {
            const int sp = a.sortedPos[a.exclList[e]];
            const int Jp = sp>>5, b = sp&31;
            if (Jp == I) {
                if (b > lane)
                    atomicOr(&out[b].y, (int) (1u<<lane));
            }
            else if (Jp > I) {
                const unsigned word = row[Jp];
                if ((word>>b)&1u)
                    atomicOr(&out[base[Jp]+__popc(word&((1u<<b)-1))].y, (int) (1u<<lane));
            }
        }
    }
}

void launchBuildList(const ListArgs& a, cudaStream_t stream, int numSMs) {
    if (a.tileFlags) {
        const int groups = (a.numBlocks+31)>>5;
        const long long warps = (long long) a.numBlocks*groups;
        k_tileFlags<<<(unsigned) ((warps+FLAG_WARPS-1)/FLAG_WARPS), FLAG_WARPS*32, 0, stream>>>(a);
        const size_t smem = sizeof(int)*(size_t) COMPACT_WARPS*a.flagStride;
        static bool attrSet = false;
        if (!attrSet) {
            cudaFuncSetAttribute(k_compactList, cudaFuncAttributeMaxDynamicSharedMemorySize, 4*COMPACT_WARPS*SNB_FLAGS_MAX_BLOCKS);
            attrSet = true;
        }
        k_compactList<<<(a.numBlocks+COMPACT_WARPS-1)/COMPACT_WARPS, COMPACT_WARPS*32, smem, stream>>>(a);
        return;
    }
    int blocks = (a.numBlocks+LIST_WARPS-1)/LIST_WARPS;
    int cap = numSMs*16;
    if (blocks > cap)
        blocks = cap;
    if (blocks < 1)
        blocks = 1;
    k_buildList<<<blocks, LIST_WARPS*32, 0, stream>>>(a);
}
