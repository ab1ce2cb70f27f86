// Sliced PME interpolation through shared-memory grid bricks (mixed precision, atoms in the sorted order of the tile list,
// large systems; every other case -- and all spreading -- keeps the kernels of pme.cu).
//
// The atoms arrive spatially sorted, so consecutive atoms and their 5x5x5 stencils sit in a small brick of the grid.  A CTA
// of 256 atoms fetches the 24x24x28 brick of the (lambda-combined) potential with ONE bulk tensor load
// (cp.async.bulk.tensor.4d, SASS UTMALDG, completion on an mbarrier) and then works one THREAD per atom: 125 shared-memory
// loads and ~460 FMAs with the atom's spline weights in registers, no shuffles, no reduction.
// A set of atoms whose stencils do not fit one brick (a jump of the space-filling curve, the periodic boundary, a subset
// boundary: the grids are per subset) is halved, in sorted order, until it does; a single atom always fits.
//
// Bricks never wrap: they address a PADDED copy of the potentials, [S][nx+4][ny+4][nzp] with nzp = nz+4 rounded up to a
// multiple of 4 floats (TMA strides are multiples of 16 bytes); a brick hanging over the padded extent is clipped by the TMA
// unit itself (the subset is the tensor's fourth coordinate).  After the inverse transform k_pmeCombinePad writes
// phi_i = sum_j lambda[slice(i,j)] grid_j straight into the padded layout, halo included.
//
// Spreading through bricks was built and measured four ways in round 2 and never beat the per-atom kernel of pme.cu
// (DESIGN.md section 7a): whole-brick bulk tensor reductions (UTMAREDG), shared-memory atomics, a warp-private 16x16x20 brick
// flushed with red.global.add.v4.f32 (256 / 48 / 22 us against 225 / 32 / 15.5 us at STMV / ApoA1 / DHFR size), and the same
// with a 12x12x16 brick for twice the resident warps (229 us at STMV size).
//
// Arithmetic: platforms/reference/src/ReferencePME.cpp:196-256 (index, fraction), :264-317 (B-splines), :598-702
// (interpolation; subset stride nx, see pme.cu).  (GPU behavioural cross-reference: platforms/common/src/kernels/pme.cc:296-408.)
#include "kernels.cuh"

#include <climits>
#include <cstdlib>

#define ORDER SNB_PME_ORDER
#define MAX_WARPS 8                      /* CTAs of 128 or 256 threads */

// Brick shape in grid points (x, y, z).  TMA wants the innermost coordinate of a box 16-byte aligned -- a start at z = 2
// raises CUDA_ERROR_ILLEGAL_INSTRUCTION, measured with tools/ubench/tma_shape_test.cu -- so a brick starts at z rounded
// down to a multiple of 4 floats and is 4 points longer along z than along x and y.
template <int BX_, int BY_, int BZ_> struct BrickShape {
    enum { BX = BX_, BY = BY_, BZ = BZ_, FLOATS = BX_*BY_*BZ_, BYTES = FLOATS*4 };
};
typedef BrickShape<24, 24, 28> BigBrick;         // 64 512 bytes

namespace {

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

// Essmann recursion for order 5 (ReferencePME.cpp:264-317); the same statements as pme.cu
__device__ __forceinline__ void bspline5f(float w, float* th, float* dth) {
    th[4] = 0.f;
    th[1] = w;
    th[0] = 1.f-w;
#pragma unroll
    for (int k = 3; k < ORDER; k++) {
        const float div = 1.f/(float) (k-1);
        th[k-1] = div*w*th[k-2];
#pragma unroll
        for (int l = 1; l < k-1; l++)
            th[k-l-1] = div*((w+l)*th[k-l-2]+(k-l-w)*th[k-l-1]);
        th[0] = div*(1.f-w)*th[0];
    }
    dth[0] = -th[0];
#pragma unroll
    for (int k = 1; k < ORDER; k++)
        dth[k] = th[k-1]-th[k];
    const float div = 1.f/(float) (ORDER-1);
    th[ORDER-1] = div*w*th[ORDER-2];
#pragma unroll
    for (int l = 1; l < ORDER-1; l++)
        th[ORDER-l-1] = div*((w+l)*th[ORDER-l-2]+(ORDER-l-w)*th[ORDER-l-1]);
    th[0] = div*(1.f-w)*th[0];
}

// this thread's atom: grid cell, fraction (fractional coordinates in double: positions may be far from the origin), charge
__device__ __forceinline__ bool threadAtom(const PmeArgs& a, int k, int& atom, int idx[3], float frac[3], float& q, int& sub) {
    atom = -1; q = 0.f; sub = -1;
    idx[0] = idx[1] = idx[2] = 0;
    frac[0] = frac[1] = frac[2] = 0.f;
    if (k >= a.numAtoms)
        return false;
    atom = a.sortedAtom[k];
    if (atom < 0)
        return false;
    // (position and parameters are requested together: one DRAM round trip after the atom's index, not two)
    const float4 prm = a.params4[atom];
    const double x = a.posd[3*(size_t) atom], y = a.posd[3*(size_t) atom+1], z = a.posd[3*(size_t) atom+2];
    q = a.term == 0 ? prm.x : 8.0f*prm.y*prm.y*prm.y*prm.z;
    if (q == 0.f)
        return false;
    sub = __float_as_int(prm.w);
    const BoxD& b = a.sp->exact.boxd;
    double t[3];
    t[0] = x*b.recip[0][0]+y*b.recip[1][0]+z*b.recip[2][0];
    t[1] = y*b.recip[1][1]+z*b.recip[2][1];
    t[2] = z*b.recip[2][2];
    const int n[3] = {a.nx, a.ny, a.nz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double u = (t[d]-floor(t[d]))*n[d];
        const int ti = (int) u;
        frac[d] = (float) (u-ti);
        idx[d] = ti%n[d];
    }
    return true;
}

// Scratch of the CTA-wide votes below.
struct SetScratch {
    int lo[MAX_WARPS][3], hi[MAX_WARPS][3];
    int count[MAX_WARPS], first[MAX_WARPS];
    int subsetOfFirst;
};

// The next set of this CTA's atoms that shares one brick: same subset, stencils within the brick along every axis.  Starts
// from all pending atoms of the first pending atom's subset and keeps the lower half (in thread = sorted order) until the
// set fits.  Every thread of the CTA calls it (it synchronises the CTA); returns whether THIS thread's atom is in the set,
// the brick's origin and its subset.
template <typename B>
__device__ __forceinline__ bool nextBrickSet(SetScratch& s, bool pending, int sub, const int idx[3], int origin[3], int& subset) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x&31, warp = threadIdx.x>>5, nwarps = blockDim.x>>5;
    // subset of the first pending atom
    {
        const unsigned m = __ballot_sync(full, pending);
        if (lane == 0)
            s.first[warp] = m ? warp*32+__ffs(m)-1 : INT_MAX;
        __syncthreads();
        int first = INT_MAX;
        for (int w = 0; w < nwarps; w++)
            first = min(first, s.first[w]);
        if ((int) threadIdx.x == first)
            s.subsetOfFirst = sub;
        __syncthreads();
        subset = s.subsetOfFirst;
    }
    const bool cand = pending && sub == subset;
    // rank of this thread among the candidates
    int rank, count;
    {
        const unsigned m = __ballot_sync(full, cand);
        if (lane == 0)
            s.count[warp] = __popc(m);
        __syncthreads();
        rank = __popc(m&((1u<<lane)-1u));
        count = 0;
        for (int w = 0; w < nwarps; w++) {
            if (w < warp)
                rank += s.count[w];
            count += s.count[w];
        }
    }
    int limit = count;
    for (;;) {
        const bool in = cand && rank < limit;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int lo = __reduce_min_sync(full, in ? idx[d] : INT_MAX), hi = __reduce_max_sync(full, in ? idx[d] : INT_MIN);
            if (lane == 0) {
                s.lo[warp][d] = lo;
                s.hi[warp][d] = hi;
            }
        }
        __syncthreads();            // (also: the count / first reads above are over before the next call rewrites them)
        bool fits = true;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            int lo = INT_MAX, hi = INT_MIN;
            for (int w = 0; w < nwarps; w++) {
                lo = min(lo, s.lo[w][d]);
                hi = max(hi, s.hi[w][d]);
            }
            origin[d] = d == 2 ? lo&~3 : lo;        // z: 16-byte aligned box start
            fits = fits && hi-origin[d]+ORDER <= (d == 0 ? (int) B::BX : d == 1 ? (int) B::BY : (int) B::BZ);
        }
        __syncthreads();            // everyone has read lo / hi
        if (fits || limit <= 1)
            return in;
        limit = (limit+1)/2;
    }
}

__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace

#define ROWS_PER_CTA 8
// ---- potentials seen by each subset, written into the padded layout (halo = periodic images) -------------------------
// One warp per padded row (x, y), all subsets: phi_k = sum_j lambda[slice(k,j)] grid_j.  VEC = 4 (nz a multiple of 4, hence
// nzp = nz+4): float4 loads and stores, the halo group [nz, nz+4) is the row's first group; VEC = 1: any nz.
template <int NS, int VEC>
__global__ void __launch_bounds__(32*ROWS_PER_CTA) k_pmeCombinePad(PmeArgs a) {
    __shared__ float sLambda[SNB_MAX_SUBSETS*SNB_MAX_SUBSETS];
    const int S = NS > 0 ? NS : a.numSubsets, tid = threadIdx.y*32+threadIdx.x;
    if (tid < S*S) {
        const int si = tid/S, sj = tid-si*S;
        sLambda[tid] = (float) (a.term == 0 ? a.sp->lambdaC : a.sp->lambdaV)[sliceOf(si, sj)];
    }
    __syncthreads();
    const int nx = a.nx, ny = a.ny, nz = a.nz, px = nx+4, py = ny+4, pz = a.nzp;
    const int row = blockIdx.x*ROWS_PER_CTA+threadIdx.y;        // x*py+y of the padded grid
    if (row >= px*py)
        return;
    const int y = row%py, x = row/py;
    const size_t P = (size_t) px*py*pz, G = (size_t) nx*ny*nz;
    const float* src = (const float*) a.grid+((size_t) (x >= nx ? x-nx : x)*ny+(y >= ny ? y-ny : y))*nz;
    float* dst = a.gridPad+(size_t) row*pz;
    constexpr int MAXS = NS > 0 ? NS : SNB_MAX_SUBSETS;
    if (VEC == 4) {
        for (int z = 4*threadIdx.x; z < pz; z += 128) {
            const int zs = z >= nz ? z-nz : z;
            float4 v[MAXS];
#pragma unroll
            for (int j = 0; j < MAXS; j++)
                if (j < S)
                    v[j] = *(const float4*) (src+j*G+zs);
#pragma unroll
            for (int k = 0; k < MAXS; k++)
                if (k < S) {
                    float4 phi = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < MAXS; j++)
                        if (j < S) {
                            const float l = sLambda[k*S+j];
                            phi.x += l*v[j].x; phi.y += l*v[j].y; phi.z += l*v[j].z; phi.w += l*v[j].w;
                        }
                    *(float4*) (dst+k*P+z) = phi;
                }
        }
    }
    else
    for (int z = threadIdx.x; z < pz; z += 32) {
        float v[MAXS];
        const bool real = z < nz+4;             // the tail of a row is alignment padding
        const int zs = z >= nz ? z-nz : z;
        for (int j = 0; j < S; j++)
            v[j] = real ? src[j*G+zs] : 0.f;
        for (int k = 0; k < S; k++) {
            float phi = 0.f;
            for (int j = 0; j < S; j++)
                phi += sLambda[k*S+j]*v[j];
            dst[k*P+z] = phi;
        }
    }
}

// ---- interpolation ----------------------------------------------------------------------------------------------------
template <typename B>
__global__ void __launch_bounds__(32*MAX_WARPS) k_pmeInterpBrick(PmeArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char sDyn[];
    const float* sBrick = (const float*) sDyn;
    __shared__ __align__(8) unsigned long long sBar;
    __shared__ SetScratch sSet;
    const unsigned bar = smemAddr(&sBar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int atom, idx[3], sub;
    float frac[3], q;
    const bool valid = threadAtom(a, blockIdx.x*blockDim.x+threadIdx.x, atom, idx, frac, q, sub);
    float th[3][ORDER], dth[3][ORDER];
#pragma unroll
    for (int d = 0; d < 3; d++)
        bspline5f(frac[d], th[d], dth[d]);
    bool pending = valid;
    unsigned phase = 0;
    while (__syncthreads_or(pending)) {         // (the barrier also covers: mbarrier initialised; everyone done with the previous brick)
        int origin[3], subset;
        const bool in = nextBrickSet<B>(sSet, pending, sub, idx, origin, subset);
        pending = pending && !in;
        if (threadIdx.x == 0) {
            fenceProxyAsync();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "n"((int) B::BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         :: "r"(smemAddr(sBrick)), "l"((unsigned long long) &tmap), "r"(bar), "r"(origin[2]), "r"(origin[1]), "r"(origin[0]), "r"(subset) : "memory");
        }
        // wait for the bytes (bounded spin: a descriptor error must not hang the GPU)
        {
            unsigned done = 0;
            for (int spin = 0; !done && spin < (1<<22); spin++)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(phase) : "memory");
            if (!done)
                __trap();
        }
        phase ^= 1u;
        if (in) {
            const float* cell = sBrick+((idx[0]-origin[0])*B::BY+idx[1]-origin[1])*B::BZ+idx[2]-origin[2];
            float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
            for (int ix = 0; ix < ORDER; ix++) {
                float tyz = 0.f, dyz = 0.f, tzd = 0.f;      // sum_y theta_y s, sum_y dtheta_y s, sum_y theta_y d
#pragma unroll
                for (int iy = 0; iy < ORDER; iy++) {
                    const float* row = cell+(ix*B::BY+iy)*B::BZ;
                    float s = 0.f, d = 0.f;
#pragma unroll
                    for (int iz = 0; iz < ORDER; iz++) {
                        const float phi = row[iz];
                        s = fmaf(th[2][iz], phi, s);
                        d = fmaf(dth[2][iz], phi, d);
                    }
                    tyz = fmaf(th[1][iy], s, tyz);
                    dyz = fmaf(dth[1][iy], s, dyz);
                    tzd = fmaf(th[1][iy], d, tzd);
                }
                fx = fmaf(dth[0][ix], tyz, fx);
                fy = fmaf(th[0][ix], dyz, fy);
                fz = fmaf(th[0][ix], tzd, fz);
            }
            // to Cartesian forces (ReferencePME.cpp:698-700), in double
            const BoxD& b = a.sp->exact.boxd;
            const double qd = q;
            const double gx = (double) fx*a.nx, gy = (double) fy*a.ny, gz = (double) fz*a.nz;
            const double Fx = -qd*(gx*b.recip[0][0]);
            const double Fy = -qd*(gx*b.recip[1][0]+gy*b.recip[1][1]);
            const double Fz = -qd*(gx*b.recip[2][0]+gy*b.recip[2][1]+gz*b.recip[2][2]);
            addFixed(&a.forceOrig[atom], toFixedD(Fx));
            addFixed(&a.forceOrig[a.paddedAtoms+atom], toFixedD(Fy));
            addFixed(&a.forceOrig[2*a.paddedAtoms+atom], toFixedD(Fz));
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
bool makeBrickTensorMap(CUtensorMap* map, float* gridPad, int numAtoms, int numSubsets, int nx, int ny, int nzp) {
    (void) numAtoms;
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status) != cudaSuccess || status != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return false;
        }
        encode = (EncodeTiled) fn;
    }
    // the padded grid is at least one brick long along every axis
    if (nx+4 < BigBrick::BX || ny+4 < BigBrick::BY || nzp < BigBrick::BZ)
        return false;
    // innermost first: z (row), y, x, subset
    const cuuint64_t dims[4] = {(cuuint64_t) nzp, (cuuint64_t) (ny+4), (cuuint64_t) (nx+4), (cuuint64_t) numSubsets};
    const cuuint64_t strides[3] = {(cuuint64_t) nzp*4, (cuuint64_t) nzp*4*(ny+4), (cuuint64_t) nzp*4*(ny+4)*(nx+4)};
    const cuuint32_t box[4] = {BigBrick::BZ, BigBrick::BY, BigBrick::BX, 1};
    const cuuint32_t elem[4] = {1, 1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gridPad, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launchPmeInterpolateBrick(const PmeArgs& a, const CUtensorMap& map, cudaStream_t stream) {
    const int rows = (a.nx+4)*(a.ny+4);
    const dim3 cgrid((rows+ROWS_PER_CTA-1)/ROWS_PER_CTA), cblock(32, ROWS_PER_CTA);
    const bool vec = a.nz%4 == 0;       // (then every row of the dense grid starts 16-byte aligned, and nzp = nz+4)
    if (vec && a.numSubsets == 1) k_pmeCombinePad<1, 4><<<cgrid, cblock, 0, stream>>>(a);
    else if (vec && a.numSubsets == 2) k_pmeCombinePad<2, 4><<<cgrid, cblock, 0, stream>>>(a);
    else if (vec && a.numSubsets <= 4) k_pmeCombinePad<4, 4><<<cgrid, cblock, 0, stream>>>(a);
    else k_pmeCombinePad<0, 1><<<cgrid, cblock, 0, stream>>>(a);
    const int threads = 32*MAX_WARPS, blocks = (a.numAtoms+threads-1)/threads;
    const cudaError_t err = cudaFuncSetAttribute(k_pmeInterpBrick<BigBrick>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) BigBrick::BYTES);
    if (err != cudaSuccess)
        return err;
    k_pmeInterpBrick<BigBrick><<<blocks, threads, BigBrick::BYTES, stream>>>(a, map);
    return cudaSuccess;
}
