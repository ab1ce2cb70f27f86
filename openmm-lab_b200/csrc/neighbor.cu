// Tile-list build (north-star subsystem 1b): one fused kernel, one CTA of four warps per block of 32 i
// atoms, one warp per octet.
//
// The sorted atoms are cut into blocks of 32 (j side) and octets of 8 (i side; 4 per block).  For
// every octet the list holds the later atoms that come within the cutoff of the octet's bounding
// box -- the tile interaction flags, compacted into chunks of 32 j atoms -- each with an 8-bit
// exclusion mask over the octet's atoms.  An octet's box is much tighter than a 32-atom block's, so
// ~40% of the pairs a tile visits are in range instead of ~27%.
//
// Candidate search is O(N): the atoms were binned into small cells (sort.cu) whose sorted ranges
// are known, so a CTA
//   1. marks, in a shared bitmap over the j blocks, the blocks owning atoms of the cells that overlap the
//      i block's box dilated by the cutoff (cells farther than the cutoff pruned; periodic wrap in
//      fractional coordinates, so triclinic boxes need nothing special) -- all 128 threads;
//   2. compacts the set bits into one ascending candidate list (all 128 threads, one prefix sum); the candidates are
//      split evenly between the four warps, each of which loads a candidate block's 32 atoms ONCE (four candidates
//      in flight) and tests them against all four octet boxes (one ballot -> one 32-bit word of interaction flags
//      per octet);
//   3. every warp then gathers its own octet's non-empty words, prefix-sums the popcounts, takes a
//      contiguous range of chunks from the global list with one atomicAdd, writes the entries (ascending
//      j: the tile kernel's gathers stay coalesced) and ORs the exclusion masks in from the octet's own
//      (short) exclusion lists.
// No order-dependent atomics: chunk CONTENTS are deterministic, only their location in the list varies.
// A rank of a multi-GPU run builds the lists of its own block range [blockLo, blockHi) only.
//
// No counterpart source exists in the reference tree (OpenMM's CudaNonbondedUtilities does this job
// there, un-vendored); exclusion semantics: every exception excludes its pair
// (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:104-112).
#include "kernels.cuh"

#include <type_traits>

#define LIST_WARPS 4
#define HIT_CAP 128            /* hit blocks buffered per batch; a fuller octet emits several batches */
#define CAND_CAP 1024          /* candidate blocks listed at a time (more: further rounds over the same bitmap) */
#define TEST_BATCH 128         /* candidates tested between two CTA barriers: 32 per warp */
#define LIST_GROUP 2           /* candidate blocks a warp loads together */

__device__ __forceinline__ float wrapOrtho(float d, float L, float invL) { return d-L*rintf(d*invL); }

__device__ __forceinline__ float boxDist2(float dx, float dy, float dz, float3 h) {
    float x = fmaxf(0.f, fabsf(dx)-h.x), y = fmaxf(0.f, fabsf(dy)-h.y), z = fmaxf(0.f, fabsf(dz)-h.z);
    return x*x+y*y+z*z;
}

// Smallest box distance over the periodic images the pair kernel can pick.
__device__ __forceinline__ float imageDist2(float dx, float dy, float dz, float3 h, const int pbcMode, const BoxF& box) {
    if (pbcMode == PBC_NONE)
        return boxDist2(dx, dy, dz, h);
    if (pbcMode != PBC_PAIR_TRICLINIC) {
        dx = wrapOrtho(dx, box.ax, box.invAx);
        dy = wrapOrtho(dy, box.by, box.invBy);
        dz = wrapOrtho(dz, box.cz, box.invCz);
        return boxDist2(dx, dy, dz, h);
    }
    // triclinic: reduce like the pair kernel does, then try the 27 neighbouring images
    float s = rintf(dz*box.invCz);
    dx -= s*box.cx; dy -= s*box.cy; dz -= s*box.cz;
    s = rintf(dy*box.invBy);
    dx -= s*box.bx; dy -= s*box.by;
    s = rintf(dx*box.invAx);
    dx -= s*box.ax;
    float best = 3.0e38f;
    for (int k = -1; k <= 1; k++)
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                float x = dx+i*box.ax+j*box.bx+k*box.cx;
                float y = dy+j*box.by+k*box.cy;
                float z = dz+k*box.cz;
                best = fminf(best, boxDist2(x, y, z, h));
            }
    return best;
}

// Range of cells [lo, lo+n) (periodic: modulo dim) along every axis that the box centre +- half can touch.
// For axis-aligned cells (orthorhombic or non-periodic) it also returns what the per-cell distance test
// needs: rel = position of cell `lo` relative to the centre, in cell units, and the cell size in nm.
__device__ __forceinline__ void cellSpan(const ListArgs& a, const double c[3], const double half[3], int lo[3], int n[3],
                                         float rel[3], float cellSize[3]) {
    if (a.periodic) {
        const double (*r)[3] = a.sp->exact.boxd.recip;
        double f[3], h[3];
        f[0] = c[0]*r[0][0]+c[1]*r[1][0]+c[2]*r[2][0];
        f[1] = c[1]*r[1][1]+c[2]*r[2][1];
        f[2] = c[2]*r[2][2];
        h[0] = half[0]*fabs(r[0][0])+half[1]*fabs(r[1][0])+half[2]*fabs(r[2][0]);
        h[1] = half[1]*fabs(r[1][1])+half[2]*fabs(r[2][1]);
        h[2] = half[2]*fabs(r[2][2]);
        for (int d = 0; d < 3; d++) {
            const int dim = a.cellDim[d];
            const int l = (int) floor((f[d]-h[d])*dim), u = (int) floor((f[d]+h[d])*dim);
            n[d] = min(u-l+1, dim);
            const int m = l%dim;
            lo[d] = m < 0 ? m+dim : m;
            rel[d] = (float) ((double) l-f[d]*dim);
            cellSize[d] = (float) (1.0/(fabs(r[d][d])*dim));
        }
    }
    else {
        for (int d = 0; d < 3; d++) {
            const int dim = a.cellDim[d];
            const double e0 = a.extent[d], span = a.extent[3+d]-e0;
            int l = 0, u = 0;
            rel[d] = 0.f;
            cellSize[d] = 0.f;
            if (span > 0) {
                l = max(0, min(dim-1, (int) floor((c[d]-half[d]-e0)/span*dim)));
                u = max(0, min(dim-1, (int) floor((c[d]+half[d]-e0)/span*dim)));
                rel[d] = (float) ((double) l-(c[d]-e0)/span*dim);
                cellSize[d] = (float) (span/dim);
            }
            lo[d] = l;
            n[d] = u-l+1;
        }
    }
}

// distance along one axis between cell `i` of the span and the octet's box, in nm
__device__ __forceinline__ float cellGap(float rel, int i, float halfCells, float cellSize) {
    const float left = rel+(float) i;                   // the cell spans [left, left+1] around the box centre
    return fmaxf(0.f, fmaxf(left-halfCells, -(left+1.f)-halfCells))*cellSize;
}

// PBC: the periodic-image rule of the launch, a template parameter: the triclinic rule tries 27 images per test, and with
// all three rules compiled into one body the kernel is ~9 500 instructions and stalls on instruction fetch (ncu: 9 % of
// the warp time with no instruction available); each copy carries only its own rule.
template <int PBC> __global__ void __launch_bounds__(LIST_WARPS*32, 7) k_buildList(ListArgs a) {
    extern __shared__ unsigned sMem[];
    const int lane = threadIdx.x&31, w = threadIdx.x>>5;
    const int I = a.blockLo+blockIdx.x;                 // i block; warp w owns its octet w
    if (I >= a.blockHi)
        return;
    // list reuse: nothing to do unless this evaluation rebuilds (sort.cu, k_listCheck)
    if (a.reuse && a.counters[SNB_COUNTER_REBUILD] == 0)
        return;
    const int O = 4*I+w;
    // shared: bitmap | candidates of the window | flag words of the batch under test | per-octet hit lists
    unsigned* bm = sMem;
    int* cand = (int*) (bm+a.bitmapWords);
    unsigned* tmpWord = (unsigned*) (cand+CAND_CAP);                // [4][TEST_BATCH]
    int* hitJ = (int*) (tmpWord+4*TEST_BATCH)+w*HIT_CAP;            // this warp's rows of [4][HIT_CAP]
    unsigned* hitWord = (unsigned*) ((int*) (tmpWord+4*TEST_BATCH)+4*HIT_CAP)+w*HIT_CAP;
    int* hitBase = (int*) (tmpWord+4*TEST_BATCH)+8*HIT_CAP+w*HIT_CAP;
    __shared__ int sWarpTotal[LIST_WARPS];
    __shared__ int sRing[LIST_WARPS][64];               // refinement: candidates waiting for a full warp of tests
    __shared__ float4 sOctAtom[LIST_WARPS][8];          // the octet's atoms relative to its centre (huge for a padding slot)
    __shared__ int sWordLo, sWordHi;            // bitmap words that got a bit: the candidate walk visits only these
    const unsigned full = 0xffffffffu;
    const int lastI = O*8+7, wFirst = I>>5;
    const float4 cI = a.blockCenter[I];
    const float cut2 = a.cutoff*a.cutoff;
    const BoxF box = a.sp->box;
    constexpr int pbcMode = PBC;
    // the four octet boxes of the block (every warp tests candidates against all four)
    float4 oc[4];
    float3 oe[4];
    bool ov[4];
#pragma unroll
    for (int o = 0; o < 4; o++) {
        oc[o] = a.octetCenter[4*I+o];
        const float4 e = a.octetExt[4*I+o];
        oe[o] = make_float3(e.x, e.y, e.z);
        ov[o] = e.x >= 0.f;                             // false: padding octet
    }
    // a candidate atom j pairs with octet o when it comes after the octet's last atom (every pair once); never with a padding octet
    int jMin[4];
#pragma unroll
    for (int o = 0; o < 4; o++)
        jMin[o] = ov[o] ? (4*I+o)*8+7 : 0x7fffffff;
    const float4 ext = a.octetExt[O];
    const float4 ocMine = a.octetCenter[O];             // (oc[w] would index the register array dynamically)
    const bool mine = ext.x >= 0.f;
    if (lane < 8) {
        const float4 p = a.posqLocal[O*8+lane];
        const bool real = a.sortedAtom[O*8+lane] >= 0;
        // a padding slot sits far away from everything
        sOctAtom[w][lane] = real ? make_float4(p.x-ocMine.x, p.y-ocMine.y, p.z-ocMine.z, 0.f) : make_float4(1e18f, 1e18f, 1e18f, 0.f);
    }
    // Orthorhombic boxes: the periodic image of a candidate atom is taken ONCE, relative to the i block's centre, when that
    // is the image every octet of the block would pick -- a box at least 2 (cutoff + block extent) + 2 block extents long;
    // smaller boxes wrap the displacement to every octet on its own.
    bool wrapOnce = false;
    if (pbcMode == PBC_SHIFT) {
        const float4 be = a.blockExt[I];
        const float reach = a.cutoff+1e-3f;
        wrapOnce = box.ax > 2.f*(reach+2.f*be.x) && box.by > 2.f*(reach+2.f*be.y) && box.cz > 2.f*(reach+2.f*be.z);
    }

    // ---- 1. candidate blocks of the whole i block (all 128 threads) -------------------------------
    for (int k = wFirst+threadIdx.x; k < a.bitmapWords; k += blockDim.x)
        bm[k] = 0u;
    if (threadIdx.x == 0) {
        sWordLo = a.bitmapWords;
        sWordHi = -1;
    }
    __syncthreads();
    if (!a.useCutoff) {
        if (threadIdx.x == 0) {
            sWordLo = wFirst;
            sWordHi = a.bitmapWords-1;
        }
        // no cutoff: every later atom interacts
        for (int k = wFirst+threadIdx.x; k < a.bitmapWords; k += blockDim.x) {
            unsigned word = full;
            if (k == wFirst) word &= full<<(I&31);
            const int rem = a.numBlocks-k*32;
            if (rem < 32) word &= rem <= 0 ? 0u : (full>>(32-rem));
            bm[k] = word;
        }
    }
    else {
        const double margin = 1e-5;     // block-local float coordinates vs the double coordinates the cells were cut from
        const float4 be = a.blockExt[I];
        const double c[3] = {(double) cI.x, (double) cI.y, (double) cI.z};
        const double half[3] = {(double) be.x+a.cutoff+margin, (double) be.y+a.cutoff+margin, (double) be.z+a.cutoff+margin};
        int lo[3], n[3];
        float rel[3], cs[3];
        cellSpan(a, c, half, lo, n, rel, cs);
        // axis-aligned cells: skip those farther than the cutoff from the block's box (the corners of the span)
        const bool prune = pbcMode != PBC_PAIR_TRICLINIC && cs[0] > 0.f && cs[1] > 0.f && cs[2] > 0.f;
        const float hc0 = prune ? (be.x+1e-5f)/cs[0] : 0.f, hc1 = prune ? (be.y+1e-5f)/cs[1] : 0.f, hc2 = prune ? (be.z+1e-5f)/cs[2] : 0.f;
        const float pruneCut2 = (a.cutoff+1e-5f)*(a.cutoff+1e-5f);
        const int nyz = n[1]*n[2], total = n[0]*nyz;
        const int dy = a.cellDim[1], dz = a.cellDim[2], dx = a.cellDim[0];
        int myLo = a.bitmapWords, myHi = -1;       // bitmap words this thread touched
        for (int base = 0; base < total; base += 2*LIST_WARPS*32) {
            int2 range[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int idx = base+u*LIST_WARPS*32+(int) threadIdx.x;
                range[u] = make_int2(0, 0);
                if (idx < total) {
                    const int ix = idx/nyz, s = idx-ix*nyz;
                    const int iy = s/n[2], iz = s-iy*n[2];
                    bool near = true;
                    if (prune) {
                        // an axis whose span was clamped to the whole box may see a cell's other image closer
                        float gx = cellGap(rel[0], ix, hc0, cs[0]), gy = cellGap(rel[1], iy, hc1, cs[1]), gz = cellGap(rel[2], iz, hc2, cs[2]);
                        if (n[0] == dx) gx = fminf(gx, cellGap(rel[0], ix+dx, hc0, cs[0]));
                        if (n[1] == dy) gy = fminf(gy, cellGap(rel[1], iy+dy, hc1, cs[1]));
                        if (n[2] == dz) gz = fminf(gz, cellGap(rel[2], iz+dz, hc2, cs[2]));
                        near = gx*gx+gy*gy+gz*gz < pruneCut2;
                    }
                    if (near) {
                        int cx = lo[0]+ix, cy = lo[1]+iy, cz = lo[2]+iz;
                        cx -= cx >= dx ? dx : 0;
                        cy -= cy >= dy ? dy : 0;
                        cz -= cz >= dz ? dz : 0;
                        range[u] = a.cellRange[(cx*dy+cy)*dz+cz];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int s0 = max(range[u].x, I*32), e0 = range[u].y;
                if (e0 > s0) {
                    for (int b = s0>>5; b <= (e0-1)>>5; b++)
                        atomicOr(&bm[b>>5], 1u<<(b&31));
                    myLo = min(myLo, s0>>10);
                    myHi = max(myHi, (e0-1)>>10);
                }
            }
        }
        myLo = __reduce_min_sync(0xffffffffu, myLo);
        myHi = __reduce_max_sync(0xffffffffu, myHi);
        if (lane == 0 && myHi >= 0) {
            atomicMin(&sWordLo, myLo);
            atomicMax(&sWordHi, myHi);
        }
    }
    __syncthreads();

    // ---- 2. + 3. test the candidates (shared between the warps), keep this octet's hits, emit batches ----
    int nh = 0;
    bool first = true, dead = false;
    // octets too wide for the one-image-per-j-atom shortcut wrap every pair instead
    // (a kept list: the octet's box may grow by skin/2 on every side, and its centre move as much, before the next rebuild)
    const int wide = (pbcMode == PBC_SHIFT && (ext.x+a.skin >= a.sp->shiftLimit[0] || ext.y+a.skin >= a.sp->shiftLimit[1] || ext.z+a.skin >= a.sp->shiftLimit[2])) ? 1 : 0;
    const int ki = O*8+(lane&7);
    const int atomK = mine ? a.sortedAtom[ki] : -1;     // atom k = lane&7 of the octet, on every lane
    const int atomI = lane < 8 ? atomK : -1;

    // Refinement of the batch's flag words: a candidate stays only if it comes within the (padded) cutoff of one of the
    // octet's ATOMS, not just of their bounding box (41 % -> 55 % of the pair slots of a chunk in range: every entry dropped here
    // saves the direct-space kernel eight pair slots).  The surviving bits are compacted through a small ring so that all 32
    // lanes test one candidate each.  Valid where the direct-space kernel takes ONE periodic image per j atom (the one nearest
    // the octet's centre): orthorhombic or no periodicity, octet not `wide`; the other octets keep the box test.
    const bool refinable = a.useCutoff && pbcMode != PBC_PAIR_TRICLINIC && !wide && a.refine;
    auto refine = [&]() {
        int* ring = sRing[w];
        int head = 0, fill = 0;
        auto testBatch = [&](int count) {
            const int e = lane < count ? ring[(head+lane)&63] : -1;
            if (e >= 0) {
                const int k = e>>5, b = e&31, J = hitJ[k];
                const float4 pj = a.posqLocal[J*32+b];
                const float4 bcJ = a.blockCenter[J];
                float dx = pj.x+(bcJ.x-cI.x)-ocMine.x, dy = pj.y+(bcJ.y-cI.y)-ocMine.y, dz = pj.z+(bcJ.z-cI.z)-ocMine.z;
                if (pbcMode != PBC_NONE) {
                    dx = wrapOrtho(dx, box.ax, box.invAx);
                    dy = wrapOrtho(dy, box.by, box.invBy);
                    dz = wrapOrtho(dz, box.cz, box.invCz);
                }
                bool keep = false;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const float4 pi = sOctAtom[w][t];
                    const float x = dx-pi.x, y = dy-pi.y, z = dz-pi.z;
                    keep |= x*x+y*y+z*z < cut2;
                }
                if (!keep)
                    atomicAnd(&hitWord[k], ~(1u<<b));
            }
            head = (head+count)&63;
            fill -= count;
        };
        for (int k = 0; k < nh; k++) {
            const unsigned word = hitWord[k];
            if ((word>>lane)&1u)
                ring[(head+fill+__popc(word&((1u<<lane)-1)))&63] = (k<<5)|lane;
            fill += __popc(word);
            __syncwarp();
            if (fill >= 32)
                testBatch(32);
        }
        if (fill > 0)
            testBatch(fill);
        __syncwarp();
    };

    auto flush = [&]() {
        if (refinable)
            refine();
        // positions of the batch's entries (the octet's own atoms occupy entries 0..7 of its first batch)
        int running = first ? 8 : 0;
        for (int k0 = 0; k0 < nh; k0 += 32) {
            const int k = k0+lane;
            const int cnt = k < nh ? __popc(hitWord[k]) : 0;
            int incl = cnt;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(full, incl, o);
                if (lane >= o) incl += t;
            }
            if (k < nh)
                hitBase[k] = running+incl-cnt;
            running += __shfl_sync(full, incl, 31);
        }
        const int total = running, n = (total+31)>>5;
        if (n == 0)
            return;
        int c0g = 0;
        if (lane == 0) {
            c0g = atomicAdd(&a.counters[0], n);
            if (c0g+n > a.maxChunks)
                a.counters[2] = 1;                      // overflow: the host grows the buffers and repeats
        }
        c0g = __shfl_sync(full, c0g, 0);
        if (c0g+n > a.maxChunks) {
            dead = true;
            return;
        }
        int2* out = a.jList+(size_t) c0g*32;
        for (int k = lane; k < n; k += 32)
            a.chunkOctet[c0g+k] = O|(wide<<30);
        // own atoms first (lower triangle masked: every pair once), then the tail padding
        if (first && lane < 8)
            out[lane] = atomI >= 0 ? make_int2(ki, (int) ((0xffu<<lane)&0xffu)) : make_int2(-1, 0xff);
        for (int k = total+lane; k < n*32; k += 32)
            out[k] = make_int2(-1, 0xff);
        __syncwarp();
        for (int k = 0; k < nh; k++) {
            const unsigned wb = hitWord[k];
            if ((wb>>lane)&1u)
                out[hitBase[k]+__popc(wb&((1u<<lane)-1))] = make_int2(hitJ[k]*32+lane, 0);
        }
        __syncwarp();
        // exclusion masks from the i side (4 lanes share an atom's list); partners before the octet are
        // that octet's business
        if (atomK >= 0) {
            const int e0 = a.exclStart[atomK], e1 = a.exclStart[atomK+1];
            const unsigned bit = 1u<<(lane&7);
            for (int e = e0+(lane>>3); e < e1; e += 4) {
                const int sp = a.sortedPos[a.exclList[e]];
                if ((sp>>3) == O) {
                    if (first && (sp&7) > (lane&7))
                        atomicOr(&out[sp&7].y, (int) bit);
                }
                else if (sp > lastI) {
                    // hit blocks are in ascending order: binary search for the partner's block
                    const int J = sp>>5, b = sp&31;
                    int l = 0, r = nh-1, found = -1;
                    while (l <= r) {
                        const int m = (l+r)>>1, v = hitJ[m];
                        if (v == J) { found = m; break; }
                        if (v < J) l = m+1; else r = m-1;
                    }
                    if (found >= 0) {
                        const unsigned word = hitWord[found];
                        if ((word>>b)&1u)
                            atomicOr(&out[hitBase[found]+__popc(word&((1u<<b)-1))].y, (int) bit);
                    }
                }
            }
        }
        __syncwarp();
        first = false;
        nh = 0;
    };

    // ---- the candidate blocks of the whole touched window [wordLo, wordHi] of the bitmap, compacted in ascending order
    // by all 128 threads at once (every thread owns a contiguous run of words; one CTA-wide prefix sum), then tested in
    // batches.  (Walking the bitmap in windows of 1 024 blocks cost four barriers per window for a handful of candidates
    // each -- the candidates of a block are scattered along the space-filling curve.)
    const int wordLo = max(wFirst, sWordLo), wordHi = sWordHi;
    const int nWords = max(0, wordHi-wordLo+1);
    const int wpt = (nWords+LIST_WARPS*32-1)/(LIST_WARPS*32);
    const int wBeg = wordLo+(int) threadIdx.x*wpt, wEnd = min(wordHi+1, wBeg+wpt);
    int myCount = 0;
    for (int k = wBeg; k < wEnd; k++)
        myCount += __popc(bm[k]);
    int myOffset, totalCand;
    {
        int incl = myCount;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31)
            sWarpTotal[w] = incl;
        __syncthreads();
        myOffset = incl-myCount;
        totalCand = 0;
        for (int k = 0; k < LIST_WARPS; k++) {
            if (k < w) myOffset += sWarpTotal[k];
            totalCand += sWarpTotal[k];
        }
    }
    for (int base = 0; base < totalCand; base += CAND_CAP) {
        // this round's slice [base, base+CAND_CAP) of the ascending candidate sequence
        if (myOffset < base+CAND_CAP && myOffset+myCount > base) {
            int pos = myOffset-base;
            for (int k = wBeg; k < wEnd; k++) {
                unsigned word = bm[k];
                while (word) {
                    if (pos >= 0 && pos < CAND_CAP)
                        cand[pos] = k*32+__ffs(word)-1;
                    pos++;
                    word &= word-1;
                }
            }
        }
        __syncthreads();
        const int nc = min(CAND_CAP, totalCand-base);
        for (int c0 = 0; c0 < nc; c0 += TEST_BATCH) {
            // the batch is split evenly between the four warps (whole groups of LIST_GROUP candidates, loaded together);
            // each tests its candidates against all four octet boxes.  An even split matters: the warps meet at a
            // barrier after the batch.  The image rule of the block (wrapOnce) is a template parameter of the loop: as a run-time
            // select BOTH rules were evaluated for every test (ncu: 225 instructions per candidate), and the loop body no
            // longer fit the instruction cache (18 % of the warp time with no instruction to issue).
            const int nbAll = min(TEST_BATCH, nc-c0);
            const int perWarp = ((nbAll+LIST_WARPS*LIST_GROUP-1)/(LIST_WARPS*LIST_GROUP))*LIST_GROUP;
            const int kBeg = c0+w*perWarp, kEnd = min(c0+nbAll, kBeg+perWarp);
            auto testCandidates = [&](auto wrapTag) {
                constexpr bool WRAP_ONCE = decltype(wrapTag)::value;
#pragma unroll 1
                for (int k0 = kBeg; k0 < kEnd; k0 += LIST_GROUP) {
                    int J[LIST_GROUP];
                    float4 pj[LIST_GROUP], bc[LIST_GROUP];
#pragma unroll
                    for (int u = 0; u < LIST_GROUP; u++) {
                        J[u] = k0+u < kEnd ? cand[k0+u] : -1;
                        pj[u] = bc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (J[u] >= 0 && a.useCutoff) {
                            pj[u] = a.posqLocal[J[u]*32+lane];
                            bc[u] = a.blockCenter[J[u]];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < LIST_GROUP; u++) {
                        if (J[u] < 0)
                            continue;
                        const int j = J[u]*32+lane;
                        const bool atom = j < a.numAtoms;
                        float px = pj[u].x+(bc[u].x-cI.x), py = pj[u].y+(bc[u].y-cI.y), pz = pj[u].z+(bc[u].z-cI.z);
                        if (WRAP_ONCE) {
                            px = wrapOrtho(px, box.ax, box.invAx);
                            py = wrapOrtho(py, box.by, box.invBy);
                            pz = wrapOrtho(pz, box.cz, box.invCz);
                        }
                        unsigned keep = 0u;
#pragma unroll
                        for (int o = 0; o < 4; o++) {
                            bool acc = atom && j > jMin[o];
                            if (a.useCutoff && acc)
                                acc = (WRAP_ONCE ? boxDist2(px-oc[o].x, py-oc[o].y, pz-oc[o].z, oe[o])
                                                 : imageDist2(px-oc[o].x, py-oc[o].y, pz-oc[o].z, oe[o], pbcMode, box)) < cut2;
                            const unsigned word = __ballot_sync(full, acc);
                            if (lane == o)
                                keep = word;
                        }
                        if (lane < 4)
                            tmpWord[lane*TEST_BATCH+(k0+u-c0)] = keep;
                    }
                }
            };
            if (pbcMode == PBC_SHIFT && wrapOnce)
                testCandidates(std::true_type());
            else
                testCandidates(std::false_type());
            __syncthreads();
            // this warp gathers its own octet's non-empty words, in candidate (= ascending j) order
            const int nb = min(TEST_BATCH, nc-c0);
            if (mine && !dead)
                for (int k0 = 0; k0 < nb; k0 += 32) {
                    if (nh > HIT_CAP-32)
                        flush();
                    if (dead)
                        break;
                    const unsigned word = k0+lane < nb ? tmpWord[w*TEST_BATCH+k0+lane] : 0u;
                    const unsigned nz = __ballot_sync(full, word != 0u);
                    if (word) {
                        const int pos = nh+__popc(nz&((1u<<lane)-1));
                        hitJ[pos] = cand[c0+k0+lane];
                        hitWord[pos] = word;
                    }
                    nh += __popc(nz);
                    __syncwarp();
                }
            __syncthreads();
        }
    }
    if (mine && !dead && (first || nh > 0))
        flush();
}

cudaError_t launchBuildList(const ListArgs& a, cudaStream_t stream, int numSMs) {
    (void) numSMs;
    const int numBlocks = a.blockHi-a.blockLo;
    if (numBlocks <= 0)
        return cudaSuccess;
    const size_t smem = sizeof(unsigned)*((size_t) a.bitmapWords+CAND_CAP+4*TEST_BATCH+12*HIT_CAP);
    // PBC_SHIFT and PBC_PAIR_ORTHO share the orthorhombic rule
    const int rule = a.pbcMode == PBC_NONE ? 0 : a.pbcMode == PBC_PAIR_TRICLINIC ? 2 : 1;
    // the attribute is per device and per context: set whenever it is needed (no process-wide cache), and checked
    if (smem > 48*1024) {
        cudaError_t err;
        if (rule == 0) err = cudaFuncSetAttribute(k_buildList<PBC_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        else if (rule == 1) err = cudaFuncSetAttribute(k_buildList<PBC_SHIFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        else err = cudaFuncSetAttribute(k_buildList<PBC_PAIR_TRICLINIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (err != cudaSuccess)
            return err;
    }
    if (rule == 0) k_buildList<PBC_NONE><<<numBlocks, LIST_WARPS*32, smem, stream>>>(a);
    else if (rule == 1) k_buildList<PBC_SHIFT><<<numBlocks, LIST_WARPS*32, smem, stream>>>(a);
    else k_buildList<PBC_PAIR_TRICLINIC><<<numBlocks, LIST_WARPS*32, smem, stream>>>(a);
    return cudaSuccess;
}
