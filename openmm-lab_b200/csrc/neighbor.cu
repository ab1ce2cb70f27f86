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
__device__ __forceinline__ void flushChunks(const ListArgs& a, int blockI, const int2* buf, int n, int lane) {
    int base = 0, item = 0;
    if (lane == 0) {
        base = atomicAdd(&a.counters[0], n);
        item = atomicAdd(&a.counters[1], 1);
        if (base+n > a.maxChunks || item >= a.maxItems)
            a.counters[2] = 1;          // overflow: the host grows the buffers and repeats the evaluation
        else
            a.workItems[item] = make_int4(blockI, base, n, 0);
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
                if (count >= SNB_ITEM_CHUNKS*32) {
                    flushChunks(a, I, buf, SNB_ITEM_CHUNKS, lane);
                    __syncwarp();
                    int2 rest = buf[SNB_ITEM_CHUNKS*32+lane];
                    __syncwarp();
                    buf[lane] = rest;
                    count -= SNB_ITEM_CHUNKS*32;
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
            flushChunks(a, I, buf, n, lane);
        __syncwarp();
    }
}

void launchBuildList(const ListArgs& a, cudaStream_t stream, int numSMs) {
    int blocks = (a.numBlocks+LIST_WARPS-1)/LIST_WARPS;
    int cap = numSMs*16;
    if (blocks > cap)
        blocks = cap;
    if (blocks < 1)
        blocks = 1;
    k_buildList<<<blocks, LIST_WARPS*32, 0, stream>>>(a);
}
