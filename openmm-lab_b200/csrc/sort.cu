// Spatial sort of the atoms (north-star subsystem 1a): bin into small cells ordered along a
// Hilbert curve, counting-sort, then cut the sorted sequence into blocks of 32 atoms with
// bounding boxes and block-local single-precision coordinates.
//
// The reference has no counterpart source for this stage: its CUDA platform leans on OpenMM's
// CudaNonbondedUtilities / ComputeContext::reorderAtoms (un-vendored; call sites
// platforms/cuda/src/CudaOpenMMLabKernels.cpp:736-758).  Written from the north-star description.
#include "snb_core.cuh"
#include "kernels.cuh"

#include <cooperative_groups.h>

// ---- 0. tile-list reuse (platform property ListSkin > 0) ---------------------------------------------
// The list of the last rebuild was built with cutoff + skin, so it holds every pair that is within the cutoff now as long
// as no atom has moved more than skin/2 since (the exact predicate of the direct-space kernel still decides the pair set
// every evaluation, so the results do not depend on when the list was built).  k_listCheck raises the rebuild flag when
// an atom has moved farther, or the host asked (first evaluation, box / capacity / partition change); without the flag the
// sort and list kernels return at once and k_blockBounds refreshes the boxes and local coordinates of the kept order.
// (OpenMM's CUDA platform, which the reference inherits its neighbour list from, reuses a padded list the same way.)
__device__ __forceinline__ bool keepList(const SortArgs& a) { return a.reuse && a.counters[SNB_COUNTER_REBUILD] == 0; }

__global__ void k_listCheck(SortArgs a) {
    const int i = blockIdx.x*blockDim.x+threadIdx.x;
    bool moved = i == 0 && a.sp->forceRebuild != 0;
    if (i < a.numAtoms) {
        const double dx = a.posd[3*i]-a.posRef[3*i], dy = a.posd[3*i+1]-a.posRef[3*i+1], dz = a.posd[3*i+2]-a.posRef[3*i+2];
        moved = moved || !(dx*dx+dy*dy+dz*dz <= a.sp->skinHalf2);      // (a NaN position rebuilds, too)
    }
    if (__any_sync(0xffffffffu, moved) && (threadIdx.x&31) == 0)
        a.counters[SNB_COUNTER_REBUILD] = 1;
}

// ---- 1. cell assignment -------------------------------------------------------------
// One thread per atom.  cellRank[] maps a linear cell index to its position along the Hilbert
// curve (host-built table).  The slot within the cell comes from an atomic counter; the
// within-cell order is made deterministic afterwards by sortCells.
__device__ __forceinline__ int cellRankOf(const SortArgs& a, double x, double y, double z) {
    int c[3];
    if (a.periodic) {
        // fractional coordinates, frac = pos * recip (lower triangular)
        double f[3];
        f[0] = x*a.sp->exact.boxd.recip[0][0]+y*a.sp->exact.boxd.recip[1][0]+z*a.sp->exact.boxd.recip[2][0];
        f[1] = y*a.sp->exact.boxd.recip[1][1]+z*a.sp->exact.boxd.recip[2][1];
        f[2] = z*a.sp->exact.boxd.recip[2][2];
        for (int d = 0; d < 3; d++) {
            double w = f[d]-floor(f[d]);
            c[d] = min(a.cellDim[d]-1, (int) (w*a.cellDim[d]));
        }
    }
    else {
        // non-periodic: the bounding box of all atoms was reduced into a.extent by k_extent
        const double* e = a.extent;
        double p[3] = {x, y, z};
        for (int d = 0; d < 3; d++) {
            double span = e[3+d]-e[d];
            double w = span > 0 ? (p[d]-e[d])/span : 0.0;
            c[d] = max(0, min(a.cellDim[d]-1, (int) (w*a.cellDim[d])));
        }
    }
    return a.cellRank[(c[0]*a.cellDim[1]+c[1])*a.cellDim[2]+c[2]];
}

__global__ void k_assignCells(SortArgs a) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (i >= a.numAtoms || keepList(a))
        return;
    double x = a.posd[3*i], y = a.posd[3*i+1], z = a.posd[3*i+2];
    if (a.reuse) {
        a.posRef[3*i] = x; a.posRef[3*i+1] = y; a.posRef[3*i+2] = z;
    }
    const int rank = cellRankOf(a, x, y, z);
    a.atomCell[i] = rank;
    a.atomSlot[i] = atomicAdd(&a.cellCount[rank], 1);
}

// bounding box of all atoms (non-periodic methods): extent[0..2] = min, [3..5] = max.
// Ordered-int trick on doubles is avoided: two-stage reduction with one block.
__global__ void k_extent(const double* posd, int numAtoms, double* extent, const int* keep) {
    if (keep && *keep == 0)
        return;
    __shared__ double smin[3][256], smax[3][256];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = threadIdx.x; i < numAtoms; i += blockDim.x)
        for (int d = 0; d < 3; d++) {
            double v = posd[3*i+d];
            lo[d] = fmin(lo[d], v);
            hi[d] = fmax(hi[d], v);
        }
    for (int d = 0; d < 3; d++) { smin[d][threadIdx.x] = lo[d]; smax[d][threadIdx.x] = hi[d]; }
    __syncthreads();
    for (int s = blockDim.x/2; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int d = 0; d < 3; d++) {
                smin[d][threadIdx.x] = fmin(smin[d][threadIdx.x], smin[d][threadIdx.x+s]);
                smax[d][threadIdx.x] = fmax(smax[d][threadIdx.x], smax[d][threadIdx.x+s]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 3) {
        extent[threadIdx.x] = smin[threadIdx.x][0];
        extent[3+threadIdx.x] = smax[threadIdx.x][0];
    }
}

// ---- 2. exclusive scan of the cell counts -----------------------------------------------------------
// Small grids (<= SCAN_SINGLE cells): one block, every thread owns a contiguous run of cells.
// Large grids: tiles of SCAN_TILE cells -- per-tile sums, a one-block scan of those, per-tile scans (coalesced).
#define SCAN_SINGLE 16384
#define SCAN_TILE 1024

__device__ __forceinline__ int blockExclusiveScan(int v, int* warpSums, int& total) {
    // exclusive prefix of v over the block (blockDim.x <= 1024); total = sum over the block
    const int lane = threadIdx.x&31, warp = threadIdx.x>>5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (int) (blockDim.x+31)/32 ? warpSums[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warpSums[lane] = w;
    }
    __syncthreads();
    total = warpSums[31];
    return (warp > 0 ? warpSums[warp-1] : 0)+incl-v;
}

// keep: null, or the rebuild flag of a handle with list reuse on (0 = nothing to do)
__global__ void __launch_bounds__(1024) k_scanCells(const int* __restrict__ count, int* __restrict__ start, int numCells, const int* keep) {
    if (keep && *keep == 0)
        return;
    __shared__ int warpSums[32];
    const int per = (numCells+blockDim.x-1)/blockDim.x;
    const int first = min(numCells, (int) threadIdx.x*per), last = min(numCells, first+per);
    int sum = 0;
    for (int i = first; i < last; i++)
        sum += count[i];
    int total;
    int running = blockExclusiveScan(sum, warpSums, total);
    for (int i = first; i < last; i++) {
        start[i] = running;
        running += count[i];
    }
    if (threadIdx.x == 0)
        start[numCells] = total;
}

__global__ void __launch_bounds__(SCAN_TILE) k_scanTileSums(const int* __restrict__ count, int* __restrict__ tileSum, int numCells, const int* keep) {
    if (keep && *keep == 0)
        return;
    __shared__ int warpSums[32];
    const int i = blockIdx.x*SCAN_TILE+threadIdx.x;
    int total;
    blockExclusiveScan(i < numCells ? count[i] : 0, warpSums, total);
    if (threadIdx.x == 0)
        tileSum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scanTileOffsets(int* __restrict__ tileSum, int numTiles, int* __restrict__ start, int numCells, const int* keep) {
    if (keep && *keep == 0)
        return;
    // exclusive scan of the tile sums in place (numTiles can exceed the block: carried loop)
    __shared__ int warpSums[32];
    int carry = 0;
    for (int base = 0; base < numTiles; base += blockDim.x) {
        const int i = base+threadIdx.x;
        const int v = i < numTiles ? tileSum[i] : 0;
        int total;
        const int excl = blockExclusiveScan(v, warpSums, total);
        if (i < numTiles)
            tileSum[i] = carry+excl;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        start[numCells] = carry;
}

__global__ void __launch_bounds__(SCAN_TILE) k_scanTiles(const int* __restrict__ count, const int* __restrict__ tileOffset, int* __restrict__ start, int numCells, const int* keep) {
    if (keep && *keep == 0)
        return;
    __shared__ int warpSums[32];
    const int i = blockIdx.x*SCAN_TILE+threadIdx.x;
    int total;
    const int excl = blockExclusiveScan(i < numCells ? count[i] : 0, warpSums, total);
    if (i < numCells)
        start[i] = tileOffset[blockIdx.x]+excl;
}

// ---- 3. scatter + deterministic order inside each cell --------------------------------
// k_scatterAtoms drops every atom into its cell's segment at the slot the atomic counter gave it
// (run-to-run order varies); k_rankInCell then gives every atom its final place = the number of
// atoms of the same cell with a smaller index, so the sorted order is a function of the positions only.
__global__ void k_scatterAtoms(SortArgs a) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (keepList(a))
        return;
    if (i < a.numAtoms)
        a.cellScratch[a.cellStart[a.atomCell[i]]+a.atomSlot[i]] = i;
    // sorted range of every cell, addressed by its linear index (the tile-list build walks cells)
    if (i < a.numCells) {
        const int r = a.cellRank[i];
        a.cellRange[i] = make_int2(a.cellStart[r], a.cellStart[r+1]);
        a.cellCount[i] = 0;         // ready for the next evaluation's k_assignCells (saves a memset per evaluation)
    }
    // padding slots of the last block
    int pad = a.numAtoms+i;
    if (i < a.paddedAtoms-a.numAtoms)
        a.sortedAtom[pad] = -1;
}

__global__ void k_rankInCell(SortArgs a) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (i >= a.numAtoms || keepList(a))
        return;
    const int c = a.atomCell[i], s = a.cellStart[c], e = a.cellStart[c+1];
    int rank = 0;
    for (int k = s; k < e; k++)
        rank += a.cellScratch[k] < i;
    a.sortedAtom[s+rank] = i;
}

// ---- 4. blocks of 32: bounding box, centre, local coordinates -------------------------
// One warp per block.  Positions are wrapped into the unit cell in double (same floor as the
// binning), the block centre is the middle of the axis-aligned bounding box, and every atom
// stores its offset from that centre in single precision: |offset| < ~1 nm, so the absolute
// coordinate error is ~3e-8 nm whatever the size of the system.
// The atom of sorted slot k, found inside its cell: the slot's rank r in the cell's segment -> the entry with exactly r
// smaller ones (entries are distinct atom indices; the scatter left them in the order of the atomic counter).  O(n^2) loads
// for a cell of n atoms, all from one or two cache lines: n is ~4 on average (periodic systems only: the cells tile the box).
__device__ __forceinline__ int atomOfSlot(const SortArgs& a, const int k) {
    const int c = a.slotCell[k], s = a.cellStart[c], e = a.cellStart[c+1], r = k-s, n = e-s;
    int atom = -1;
    if (n <= 8) {
        // the usual case: the cell's entries in registers (eight independent loads in flight), 64 compares
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            v[j] = j < n ? a.cellScratch[s+j] : 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int smaller = 0;
#pragma unroll
            for (int q = 0; q < 8; q++)
                smaller += v[q] < v[j];
            if (smaller == r && j < n)
                atom = v[j];
        }
        return atom;
    }
    for (int m = s; m < e; m++) {
        const int v = a.cellScratch[m];
        int smaller = 0;
        for (int m2 = s; m2 < e; m2++)
            smaller += a.cellScratch[m2] < v;
        if (smaller == r)
            atom = v;
    }
    return atom;
}

// RANK: the sorted order is settled here (k_boundsRank, k_sortFused) instead of by a k_rankInCell pass before
template <bool RANK> __device__ __forceinline__ void blockBoundsWarp(const SortArgs& a, const int warp, const int lane) {
    int k = warp*SNB_BLOCK+lane;
    int atom;
    if (RANK) {
        atom = k < a.numAtoms ? atomOfSlot(a, k) : -1;
        a.sortedAtom[k] = atom;
    }
    else
        atom = a.sortedAtom[k];
    double p[3] = {0, 0, 0};
    float4 prm = make_float4(0.f, 0.f, 0.f, 0.f);
    if (atom >= 0) {
        p[0] = a.posd[3*atom]; p[1] = a.posd[3*atom+1]; p[2] = a.posd[3*atom+2];
        if (a.periodic) {
            double fz, fy, fx;
            if (keepList(a)) {
                // the kept order was cut from atoms wrapped by THESE box vectors: an atom that has crossed a face of the unit
                // cell since stays with its block (a fresh floor would send it to the opposite face)
                const float4 w = a.wrapImage[atom];
                fx = w.x; fy = w.y; fz = w.z;
            }
            else {
                fz = floor(p[2]*a.sp->exact.boxd.recip[2][2]);
                fy = floor(p[1]*a.sp->exact.boxd.recip[1][1]+p[2]*a.sp->exact.boxd.recip[2][1]);
                fx = floor(p[0]*a.sp->exact.boxd.recip[0][0]+p[1]*a.sp->exact.boxd.recip[1][0]+p[2]*a.sp->exact.boxd.recip[2][0]);
                if (a.reuse)
                    a.wrapImage[atom] = make_float4((float) fx, (float) fy, (float) fz, 0.f);
            }
            p[0] -= fx*a.sp->exact.boxd.ax+fy*a.sp->exact.boxd.bx+fz*a.sp->exact.boxd.cx;
            p[1] -= fy*a.sp->exact.boxd.by+fz*a.sp->exact.boxd.cy;
            p[2] -= fz*a.sp->exact.boxd.cz;
        }
        prm = a.params4[atom];
        a.sortedPos[atom] = k;
    }
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
        lo[d] = atom >= 0 ? p[d] : 1e300;
        hi[d] = atom >= 0 ? p[d] : -1e300;
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    }
    // the centre is snapped to a 2^-10 nm lattice and rounded to float: every centre, and every difference of
    // two centres (up to 8 um apart), is then EXACT in single precision, so the kernels shift a j atom into
    // an i block's frame with float arithmetic and lose nothing
    double c[3];
    for (int d = 0; d < 3; d++)
        c[d] = (double) (float) (rint(0.5*(lo[d]+hi[d])*1024.0)*(1.0/1024.0));
    float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (atom >= 0)
        pl = make_float4((float) (p[0]-c[0]), (float) (p[1]-c[1]), (float) (p[2]-c[2]), prm.x*a.sqrtK);
    a.posqLocal[k] = pl;
    float sig = prm.y, eps = prm.z;
    // (sigma/2, 2 sqrt(eps), subset as an exact small real with -1 = padding lane, c6 = 8 s^3 e)
    const float subsetF = atom >= 0 ? (float) __float_as_int(prm.w) : -1.0f;
    a.sparams[k] = make_float4(sig, eps, subsetF, 8.0f*sig*sig*sig*eps);
    if (a.posqLocalD) {
        double4 pd = make_double4(0, 0, 0, 0), sd = make_double4(0, 0, -1, 0);
        if (atom >= 0) {
            const double sg = a.paramsD[3*atom], ep = a.paramsD[3*atom+1], qd = a.paramsD[3*atom+2];
            pd = make_double4(p[0]-c[0], p[1]-c[1], p[2]-c[2], qd*sqrt(SNB_ONE_4PI_EPS0));
            sd = make_double4(sg, ep, (double) subsetF, 8.0*sg*sg*sg*ep);
        }
        a.posqLocalD[k] = pd;
        a.sparamsD[k] = sd;
    }
    if (lane == 0) {
        a.blockCenter[warp] = make_float4((float) c[0], (float) c[1], (float) c[2], 0.f);
        // half extents, rounded up so the box is conservative in single precision
        const float4 ext = make_float4(__double2float_ru(0.5*(hi[0]-lo[0])), __double2float_ru(0.5*(hi[1]-lo[1])),
                                       __double2float_ru(0.5*(hi[2]-lo[2])), 0.f);
        a.blockExt[warp] = ext;
        // the direct-space kernels widen their cutoff band when some block is large (block-local coordinates lose bits)
        const float widest = fmaxf(ext.x, fmaxf(ext.y, ext.z));
        if (widest > SNB_NARROW_MAX_EXTENT)
            atomicMax(&a.counters[SNB_COUNTER_MAX_EXTENT], __float_as_int(widest));
    }
    // bounding boxes of the four octets (8 consecutive atoms) of the block, in block-local coordinates:
    // the i side of the tile list works on octets (a tighter box => fewer out-of-range pairs per tile)
    float ol[3] = {pl.x, pl.y, pl.z}, oh[3] = {pl.x, pl.y, pl.z};
    if (atom < 0)
        for (int d = 0; d < 3; d++) { ol[d] = 3.0e38f; oh[d] = -3.0e38f; }
    for (int d = 0; d < 3; d++)
        for (int o = 4; o > 0; o >>= 1) {
            ol[d] = fminf(ol[d], __shfl_xor_sync(0xffffffffu, ol[d], o));
            oh[d] = fmaxf(oh[d], __shfl_xor_sync(0xffffffffu, oh[d], o));
        }
    if ((lane&7) == 0) {
        const int oct = warp*4+(lane>>3);
        if (ol[0] > oh[0]) {                      // empty octet (padding)
            a.octetCenter[oct] = make_float4(0.f, 0.f, 0.f, 0.f);
            a.octetExt[oct] = make_float4(-1.f, -1.f, -1.f, 0.f);
        }
        else {
            const float cx = 0.5f*(ol[0]+oh[0]), cy = 0.5f*(ol[1]+oh[1]), cz = 0.5f*(ol[2]+oh[2]);
            a.octetCenter[oct] = make_float4(cx, cy, cz, 0.f);
            // max distance from the rounded centre to either face, plus an ulp of slack
            a.octetExt[oct] = make_float4(fmaxf(oh[0]-cx, cx-ol[0])*1.000001f+1e-7f, fmaxf(oh[1]-cy, cy-ol[1])*1.000001f+1e-7f,
                                          fmaxf(oh[2]-cz, cz-ol[2])*1.000001f+1e-7f, 0.f);
        }
    }
}

__global__ void k_blockBounds(SortArgs a) {
    const int warp = (blockIdx.x*blockDim.x+threadIdx.x)>>5, lane = threadIdx.x&31;
    if (warp < a.numBlocks)
        blockBoundsWarp<false>(a, warp, lane);
}

// ---- 5. the whole sort as ONE cooperative kernel (small and medium systems) ---------------------------------
// Below ~100 000 atoms every kernel above runs for 3-5 us of which most is launch ramp and one chain of dependent memory
// accesses, and the six of them (with the position conversion) are the head of BOTH critical chains of an evaluation
// (list -> direct space, and reciprocal space).  Here they are the phases of one kernel, separated by two grid-wide barriers
// (cooperative launch: every CTA is resident), and the exclusive scan of the cell counts is done by every CTA for itself
// in shared memory (at most SORT_FUSED_MAX_CELLS cells), which saves the barrier around a one-CTA scan:
//   A  cell of every atom, slot by atomic counter                                                     | barrier
//   B  scan of the counts -> shared memory; C  scatter into the cell segments, sorted range of every cell | barrier
//   D  the atom of every sorted slot (rank inside its cell: deterministic order); E  blocks of 32: boxes, centres, local
//      coordinates (blockBoundsWarp); counters re-zeroed.
// Periodic methods without list reuse; everything else takes the separate kernels.
#define SORT_FUSED_THREADS 256

// Grid-wide barrier of the fused kernel: one arrive counter (counters[SNB_COUNTER_BARRIER], zeroed with the rest of the
// evaluation's accumulators before the kernel starts), barrier number n is passed when the count reaches n * gridDim.x.
// The cooperative launch guarantees that every CTA is resident, so the spin cannot starve one.  (cooperative_groups'
// grid.sync() measured 4-5 us per barrier here; this one is an atomic and a polled line in the L2.)
__device__ __forceinline__ void gridBarrier(int* counter, int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1);
        int seen;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

// Exclusive scan of the cell counts into shared memory, by the whole CTA (sStart[numCells] = total): the counts arrive
// coalesced, every thread then owns a contiguous run of cells.
__device__ __forceinline__ void scanCountsShared(const int* cellCount, const int numCells, int* sStart, int* warpSums) {
    // (loads first, stores after, four int4 per thread in flight: issued one at a time behind their own shared-memory
    // stores the 23 loads of a DHFR-size grid were 14 us of latency -- ncu, gpurun r3e)
    const int n4 = numCells>>2;
    const int4* c4 = (const int4*) cellCount;
    int4* s4 = (int4*) sStart;
    for (int base = 0; base < n4; base += 4*(int) blockDim.x) {
        int4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base+u*(int) blockDim.x+(int) threadIdx.x;
            v[u] = idx < n4 ? __ldcg(c4+idx) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base+u*(int) blockDim.x+(int) threadIdx.x;
            if (idx < n4)
                s4[idx] = v[u];
        }
    }
    for (int i = 4*n4+(int) threadIdx.x; i < numCells; i += blockDim.x)
        sStart[i] = __ldcg(cellCount+i);
    __syncthreads();
    const int per = (numCells+(int) blockDim.x-1)/(int) blockDim.x;
    const int first = min(numCells, (int) threadIdx.x*per), last = min(numCells, first+per);
    int sum = 0;
#pragma unroll 8
    for (int i = first; i < last; i++)
        sum += sStart[i];
    int total;
    int running = blockExclusiveScan(sum, warpSums, total);
#pragma unroll 8
    for (int i = first; i < last; i++) {
        const int v = sStart[i];
        sStart[i] = running;
        running += v;
    }
    if (threadIdx.x == 0)
        sStart[numCells] = total;
    __syncthreads();
}

__global__ void __launch_bounds__(SORT_FUSED_THREADS) k_sortFused(SortArgs a) {
    extern __shared__ __align__(16) int sStart[];       // [numCells+1] exclusive scan of the cell counts
    __shared__ int warpSums[32];
    int* const barrier = a.counters+SNB_COUNTER_BARRIER;
    const int numCtas = (int) gridDim.x;
    const int tid = blockIdx.x*blockDim.x+threadIdx.x, nThreads = gridDim.x*blockDim.x;
    // ---- A
    for (int t = tid; t < a.numAtoms; t += nThreads) {
        const int i = t;
        const double x = a.posd[3*i], y = a.posd[3*i+1], z = a.posd[3*i+2];
        const int rank = cellRankOf(a, x, y, z);
        a.atomCell[i] = rank;
        a.atomSlot[i] = atomicAdd(&a.cellCount[rank], 1);
    }
    gridBarrier(barrier, 1*numCtas);
    // ---- B (every CTA for itself) + C
    scanCountsShared(a.cellCount, a.numCells, sStart, warpSums);
    for (int t = tid; t < a.numAtoms; t += nThreads) {
        const int c = __ldcg(a.atomCell+t), pos = sStart[c]+__ldcg(a.atomSlot+t);
        a.cellScratch[pos] = t;
        a.slotCell[pos] = c;
    }
    for (int i = tid; i < a.numCells; i += nThreads) {
        const int r = a.cellRank[i];
        a.cellRange[i] = make_int2(sStart[r], sStart[r+1]);
    }
    for (int i = tid; i <= a.numCells; i += nThreads)
        a.cellStart[i] = sStart[i];
    gridBarrier(barrier, 2*numCtas);
    // ---- D + E (every CTA is past its scan: the counters can be re-zeroed for the next evaluation)
    for (int i = tid; i < a.numCells; i += nThreads)
        a.cellCount[i] = 0;
    const int lane = threadIdx.x&31;
    for (int warp = tid>>5; warp < a.numBlocks; warp += nThreads>>5)
        blockBoundsWarp<true>(a, warp, lane);
}

// ---- 6. the same three phases as three kernels (periodic systems of up to SORT_SMEM_MAX_CELLS cells, no list reuse):
// k_assignCells -> k_scanScatter (scan in every CTA's shared memory + scatter) -> k_boundsRank (rank + block bounds),
// against five with k_scanCells and k_rankInCell as kernels of their own.
__global__ void __launch_bounds__(SORT_FUSED_THREADS) k_scanScatter(SortArgs a) {
    extern __shared__ __align__(16) int sStart[];
    __shared__ int warpSums[32];
    const int tid = blockIdx.x*blockDim.x+threadIdx.x, nThreads = gridDim.x*blockDim.x;
    scanCountsShared(a.cellCount, a.numCells, sStart, warpSums);
    for (int t = tid; t < a.numAtoms; t += nThreads) {
        const int c = a.atomCell[t], pos = sStart[c]+a.atomSlot[t];
        a.cellScratch[pos] = t;
        a.slotCell[pos] = c;
    }
    for (int i = tid; i < a.numCells; i += nThreads) {
        const int r = a.cellRank[i];
        a.cellRange[i] = make_int2(sStart[r], sStart[r+1]);
    }
    for (int i = tid; i <= a.numCells; i += nThreads)
        a.cellStart[i] = sStart[i];
}

__global__ void __launch_bounds__(SORT_FUSED_THREADS) k_boundsRank(SortArgs a) {
    const int tid = blockIdx.x*blockDim.x+threadIdx.x, nThreads = gridDim.x*blockDim.x;
    for (int i = tid; i < a.numCells; i += nThreads)
        a.cellCount[i] = 0;
    const int lane = threadIdx.x&31;
    for (int warp = tid>>5; warp < a.numBlocks; warp += nThreads>>5)
        blockBoundsWarp<true>(a, warp, lane);
}

static bool sortCompactApplies(const SortArgs& a) {
    return a.compact && !a.reuse && a.periodic && a.numCells <= SORT_SMEM_MAX_CELLS && a.numAtoms > 0;
}

// Grid of the fused kernel if this is its case and it fits on the device (every CTA resident), else 0: the separate kernels.
int sortFusedGrid(const SortArgs& a) {
    if (!a.fused || a.reuse || !a.periodic || a.numCells > SORT_FUSED_MAX_CELLS || a.numAtoms <= 0)
        return 0;
    const size_t smem = sizeof(int)*((size_t) a.numCells+1);
    int perSM = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sortFused, SORT_FUSED_THREADS, smem) != cudaSuccess || perSM < 1) {
        cudaGetLastError();
        return 0;
    }
    const int want = (a.paddedAtoms+SORT_FUSED_THREADS-1)/SORT_FUSED_THREADS;
    return max(1, min(want, perSM*max(a.numSMs, 1)));
}

static bool launchSortFused(const SortArgs& a, cudaStream_t stream, cudaError_t& err) {
    err = cudaSuccess;
    const int grid = sortFusedGrid(a);
    if (grid <= 0)
        return false;
    const size_t smem = sizeof(int)*((size_t) a.numCells+1);
    SortArgs copy = a;
    void* args[] = {&copy};
    err = cudaLaunchCooperativeKernel((const void*) k_sortFused, dim3(grid), dim3(SORT_FUSED_THREADS), args, smem, stream);
    return true;
}

cudaError_t launchSort(const SortArgs& a, cudaStream_t stream, bool* usedFused) {
    cudaError_t fusedErr;
    const bool fused = launchSortFused(a, stream, fusedErr);
    if (usedFused)
        *usedFused = fused;
    if (fused)
        return fusedErr;
    if (sortCompactApplies(a)) {
        const int t = SORT_FUSED_THREADS;
        const size_t smem = sizeof(int)*((size_t) a.numCells+1);
        k_assignCells<<<(a.numAtoms+t-1)/t, t, 0, stream>>>(a);
        k_scanScatter<<<(max(a.numAtoms, a.numCells+1)+t-1)/t, t, smem, stream>>>(a);
        k_boundsRank<<<(a.numBlocks*32+t-1)/t, t, 0, stream>>>(a);
        return cudaGetLastError();
    }
    int threads = 256;
    int atomBlocks = (a.paddedAtoms+threads-1)/threads;
    const int* keep = a.reuse ? a.counters+SNB_COUNTER_REBUILD : nullptr;
    if (a.reuse)
        k_listCheck<<<atomBlocks, threads, 0, stream>>>(a);
    if (!a.periodic)
        k_extent<<<1, 256, 0, stream>>>(a.posd, a.numAtoms, a.extent, keep);
    k_assignCells<<<atomBlocks, threads, 0, stream>>>(a);
    if (a.numCells <= SCAN_SINGLE)
        k_scanCells<<<1, 1024, 0, stream>>>(a.cellCount, a.cellStart, a.numCells, keep);
    else {
        const int tiles = (a.numCells+SCAN_TILE-1)/SCAN_TILE;
        k_scanTileSums<<<tiles, SCAN_TILE, 0, stream>>>(a.cellCount, a.tileSum, a.numCells, keep);
        k_scanTileOffsets<<<1, 1024, 0, stream>>>(a.tileSum, tiles, a.cellStart, a.numCells, keep);
        k_scanTiles<<<tiles, SCAN_TILE, 0, stream>>>(a.cellCount, a.tileSum, a.cellStart, a.numCells, keep);
    }
    k_scatterAtoms<<<(max(a.paddedAtoms, a.numCells)+threads-1)/threads, threads, 0, stream>>>(a);
    k_rankInCell<<<atomBlocks, threads, 0, stream>>>(a);
    k_blockBounds<<<(a.numBlocks*32+threads-1)/threads, threads, 0, stream>>>(a);
    return cudaSuccess;
}
