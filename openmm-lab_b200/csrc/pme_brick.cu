// Sliced PME spreading and interpolation through shared-memory grid bricks moved by TMA (mixed precision, atoms in
// the sorted order of the tile list; every other case keeps the kernels of pme.cu).
//
// The atoms arrive spatially sorted, so 32 consecutive atoms and their 5x5x5 stencils sit in a small brick of the
// grid.  A one-warp CTA takes 32 atoms and
//   spreading      accumulates their charges into a 16x16x20 brick in shared memory -- one atom at a time, lane ->
//                  (iy, iz) of the stencil's yz face, serial over x: plain read-modify-write, no atomics, a fixed order --
//                  and adds the brick to the grid with ONE bulk tensor reduction (cp.reduce.async.bulk.tensor.3d ... .add,
//                  SASS UTMAREDG) instead of 125 scattered atomics per atom;
//   interpolation  fetches the brick of the (lambda-combined) potential with ONE bulk tensor load (cp.async.bulk.tensor.3d,
//                  SASS UTMALDG, completion on an mbarrier) and then works one THREAD per atom: 125 shared-memory loads
//                  and ~460 FMAs with the atom's spline weights in registers, no shuffles, no reduction.
// A set of atoms whose stencils do not fit one brick (a jump of the space-filling curve, a subset boundary: the grids are
// per subset) is halved until it does; a single atom always fits.
//
// Bricks never wrap: they address a PADDED copy of the grids, [S][nx+4][ny+4][nzp] with nzp = nz+4 rounded up to a multiple
// of 4 floats (TMA strides are multiples of 16 bytes).  Spreading accumulates into it and k_pmeFold folds the halo back
// while writing the dense grid cuFFT reads (no memset of that grid, no atomics); after the inverse transform
// k_pmeCombinePad writes phi_i = sum_j lambda[slice(i,j)] grid_j straight into the padded layout, halo included.
//
// Arithmetic: platforms/reference/src/ReferencePME.cpp:196-256 (index, fraction), :264-317 (B-splines), :320-396 (spreading),
// :598-702 (interpolation; subset stride nx, see pme.cu).  (GPU behavioural cross-reference:
// platforms/common/src/kernels/pme.cc:27-135,296-408.)
#include "kernels.cuh"

#include <climits>

#define ORDER SNB_PME_ORDER
// Brick shape in grid points (x, y, z).  TMA wants the innermost coordinate of a box 16-byte aligned -- a start at z = 2
// raises CUDA_ERROR_ILLEGAL_INSTRUCTION, measured with tools/ubench/tma_shape_test.cu -- so a brick starts at z rounded
// down to a multiple of 4 floats and is 4 points longer along z to keep 16 usable ones whatever the alignment.
#define BRICK 16                         /* x and y: stencil bases may span BRICK-ORDER+1 = 12 cells */
#define BRICK_Z 20
#define BRICK_FLOATS (BRICK*BRICK*BRICK_Z)
#define BRICK_BYTES (BRICK_FLOATS*4)

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

// this lane's atom: grid cell, fraction (fractional coordinates in double: positions may be far from the origin), charge
__device__ __forceinline__ bool laneAtom(const PmeArgs& a, int k, int& atom, int idx[3], float frac[3], float& q, int& sub) {
    atom = -1; q = 0.f; sub = -1;
    idx[0] = idx[1] = idx[2] = 0;
    frac[0] = frac[1] = frac[2] = 0.f;
    if (k >= a.numAtoms)
        return false;
    atom = a.sortedAtom[k];
    if (atom < 0)
        return false;
    const float4 prm = a.params4[atom];
    q = a.term == 0 ? prm.x : 8.0f*prm.y*prm.y*prm.y*prm.z;
    if (q == 0.f)
        return false;
    sub = __float_as_int(prm.w);
    const double x = a.posd[3*(size_t) atom], y = a.posd[3*(size_t) atom+1], z = a.posd[3*(size_t) atom+2];
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

// The next set of atoms that shares one brick: same subset, stencils within BRICK points along every axis.  Starts from
// all pending atoms of the first pending atom's subset and keeps the lower half (in lane = sorted order) until the set fits.
__device__ __forceinline__ unsigned nextBrickSet(unsigned pending, int lane, int sub, const int idx[3], int origin[3], int& subset) {
    const unsigned full = 0xffffffffu;
    subset = __shfl_sync(full, sub, __ffs(pending)-1);
    unsigned cand = __ballot_sync(full, ((pending>>lane)&1u) && sub == subset);
    for (;;) {
        const bool in = (cand>>lane)&1u;
        bool fits = true;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int lo = __reduce_min_sync(full, in ? idx[d] : INT_MAX), hi = __reduce_max_sync(full, in ? idx[d] : INT_MIN);
            origin[d] = d == 2 ? lo&~3 : lo;        // z: 16-byte aligned box start
            fits = fits && hi-origin[d]+ORDER <= (d == 2 ? BRICK_Z : BRICK);
        }
        const int count = __popc(cand);
        if (fits || count <= 1)
            return cand;
        const int rank = __popc(cand&((1u<<lane)-1u));
        cand = __ballot_sync(full, in && rank < count/2);
    }
}

__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace

// ---- spreading ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_pmeSpreadBrick(PmeArgs a, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) float sBrick[BRICK_FLOATS];
    __shared__ float sTh[3][ORDER][32];         // spline weights of the 32 atoms (theta_x carries the charge)
    __shared__ int sRel[3][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    int atom, idx[3], sub;
    float frac[3], q;
    const bool valid = laneAtom(a, blockIdx.x*32+lane, atom, idx, frac, q, sub);
    {
        float th[ORDER], dth[ORDER];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            bspline5f(frac[d], th, dth);
#pragma unroll
            for (int i = 0; i < ORDER; i++)
                sTh[d][i][lane] = d == 0 ? q*th[i] : th[i];
        }
    }
    const int iy = lane < ORDER*ORDER ? lane/ORDER : 0, iz = lane < ORDER*ORDER ? lane-iy*ORDER : 0;
    unsigned pending = __ballot_sync(full, valid);
    bool issued = false;
    while (pending) {
        int origin[3], subset;
        const unsigned set = nextBrickSet(pending, lane, sub, idx, origin, subset);
        pending &= ~set;
        // the previous brick must have been READ by the reduction before it is cleared
        if (issued) {
            if (lane == 0)
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < BRICK_FLOATS/(32*4); k++)
            ((float4*) sBrick)[k*32+lane] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int d = 0; d < 3; d++)
            sRel[d][lane] = idx[d]-origin[d];
        __syncwarp();
        // one atom at a time (the next atom may touch the same cells from other lanes); the NEXT atom's weights and cell
        // are fetched from shared memory before the current one's read-modify-write, so that the two latencies overlap
        auto fetch = [&](int t, int& cell, float& wyz, float (&wx)[ORDER]) {
            cell = (sRel[0][t]*BRICK+sRel[1][t]+iy)*BRICK_Z+sRel[2][t]+iz;
            wyz = sTh[1][iy][t]*sTh[2][iz][t];
#pragma unroll
            for (int ix = 0; ix < ORDER; ix++)
                wx[ix] = sTh[0][ix][t];
        };
        unsigned m = set;
        int cell;
        float wyz, wx[ORDER];
        fetch(__ffs(m)-1, cell, wyz, wx);
        m &= m-1u;
        for (;;) {
            int cellN = 0;
            float wyzN = 0.f, wxN[ORDER] = {0.f, 0.f, 0.f, 0.f, 0.f};
            const bool more = m != 0u;
            if (more)
                fetch(__ffs(m)-1, cellN, wyzN, wxN);
            if (lane < ORDER*ORDER) {
#pragma unroll
                for (int ix = 0; ix < ORDER; ix++)
                    sBrick[cell+ix*BRICK*BRICK_Z] += wyz*wx[ix];
            }
            __syncwarp();
            if (!more)
                break;
            m &= m-1u;
            cell = cellN; wyz = wyzN;
#pragma unroll
            for (int ix = 0; ix < ORDER; ix++)
                wx[ix] = wxN[ix];
        }
        // shared-memory writes of the generic proxy -> visible to the bulk-copy engine; then brick += into the padded grid
        fenceProxyAsync();
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
                         :: "l"((unsigned long long) &tmap), "r"(smemAddr(sBrick)), "r"(origin[2]), "r"(origin[1]), "r"(subset*(a.nx+4)+origin[0]) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        issued = true;
    }
    // the CTA may leave once the engine has READ the brick; the writes complete on their own before the grid does
    if (issued && lane == 0)
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- halo fold: dense grid (cuFFT input) from the padded accumulation grid --------------------------------------------
// One warp per grid row (s, x, y): lanes run along z, so the index arithmetic is per row and every access is coalesced.
#define ROWS_PER_CTA 8
__global__ void __launch_bounds__(32*ROWS_PER_CTA) k_pmeFold(PmeArgs a) {
    const int nx = a.nx, ny = a.ny, nz = a.nz, py = ny+4, pz = a.nzp;
    const int row = blockIdx.x*ROWS_PER_CTA+threadIdx.y;        // (s*nx+x)*ny+y
    if (row >= a.numSubsets*nx*ny)
        return;
    const int y = row%ny, sx = row/ny, x = sx%nx, s = sx/nx;
    const float* base = a.gridPad+(size_t) s*(nx+4)*py*pz;
    float* out = (float*) a.grid+(size_t) row*nz;
    // the row itself and its images in the halo (the halo is 4 points wide and every dimension has at least 16)
    const float* src[4];
    int ns = 0;
    for (int ax = 0; ax <= (x < 4 ? 1 : 0); ax++)
        for (int ay = 0; ay <= (y < 4 ? 1 : 0); ay++)
            src[ns++] = base+((size_t) (x+ax*nx)*py+(y+ay*ny))*pz;
    for (int z = threadIdx.x; z < nz; z += 32) {
        float sum = 0.f;
        for (int k = 0; k < ns; k++) {
            sum += src[k][z];
            if (z < 4)
                sum += src[k][z+nz];
        }
        out[z] = sum;
    }
}

// ---- potentials seen by each subset, written into the padded layout (halo = periodic images) -------------------------
// One warp per padded row (x, y), all subsets: phi_k = sum_j lambda[slice(k,j)] grid_j.
__global__ void __launch_bounds__(32*ROWS_PER_CTA) k_pmeCombinePad(PmeArgs a) {
    __shared__ float sLambda[SNB_MAX_SUBSETS*SNB_MAX_SUBSETS];
    const int S = a.numSubsets, tid = threadIdx.y*32+threadIdx.x;
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
    for (int z = threadIdx.x; z < pz; z += 32) {
        float v[SNB_MAX_SUBSETS];
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
__global__ void __launch_bounds__(32) k_pmeInterpBrick(PmeArgs a, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) float sBrick[BRICK_FLOATS];
    __shared__ __align__(8) unsigned long long sBar;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const unsigned bar = smemAddr(&sBar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    int atom, idx[3], sub;
    float frac[3], q;
    const bool valid = laneAtom(a, blockIdx.x*32+lane, atom, idx, frac, q, sub);
    float th[3][ORDER], dth[3][ORDER];
#pragma unroll
    for (int d = 0; d < 3; d++)
        bspline5f(frac[d], th[d], dth[d]);
    unsigned pending = __ballot_sync(full, valid);
    unsigned phase = 0;
    while (pending) {
        int origin[3], subset;
        const unsigned set = nextBrickSet(pending, lane, sub, idx, origin, subset);
        pending &= ~set;
        // every lane is done with the previous brick (__ballot_sync inside nextBrickSet ordered the warp)
        fenceProxyAsync();
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "n"(BRICK_BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         :: "r"(smemAddr(sBrick)), "l"((unsigned long long) &tmap), "r"(bar), "r"(origin[2]), "r"(origin[1]), "r"(subset*(a.nx+4)+origin[0]) : "memory");
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
        if ((set>>lane)&1u) {
            const float* cell = sBrick+((idx[0]-origin[0])*BRICK+idx[1]-origin[1])*BRICK_Z+idx[2]-origin[2];
            float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
            for (int ix = 0; ix < ORDER; ix++) {
                float tyz = 0.f, dyz = 0.f, tzd = 0.f;      // sum_y theta_y s, sum_y dtheta_y s, sum_y theta_y d
#pragma unroll
                for (int iy = 0; iy < ORDER; iy++) {
                    const float* row = cell+(ix*BRICK+iy)*BRICK_Z;
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
        __syncwarp();
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
bool makeBrickTensorMap(CUtensorMap* map, float* gridPad, int numSubsets, int nx, int ny, int nzp) {
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
    // innermost first: z (row), y, then x with the subsets stacked along it
    const cuuint64_t dims[3] = {(cuuint64_t) nzp, (cuuint64_t) (ny+4), (cuuint64_t) numSubsets*(nx+4)};
    const cuuint64_t strides[2] = {(cuuint64_t) nzp*4, (cuuint64_t) nzp*4*(ny+4)};
    const cuuint32_t box[3] = {BRICK_Z, BRICK, BRICK}, elem[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, gridPad, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void launchPmeSpreadBrick(const PmeArgs& a, const CUtensorMap& map, cudaStream_t stream) {
    const size_t padFloats = (size_t) a.numSubsets*(a.nx+4)*(a.ny+4)*a.nzp;
    cudaMemsetAsync(a.gridPad, 0, sizeof(float)*padFloats, stream);
    k_pmeSpreadBrick<<<(a.numAtoms+31)/32, 32, 0, stream>>>(a, map);
    const int rows = a.numSubsets*a.nx*a.ny;
    k_pmeFold<<<(rows+ROWS_PER_CTA-1)/ROWS_PER_CTA, dim3(32, ROWS_PER_CTA), 0, stream>>>(a);
}

void launchPmeInterpolateBrick(const PmeArgs& a, const CUtensorMap& map, cudaStream_t stream) {
    const int rows = (a.nx+4)*(a.ny+4);
    k_pmeCombinePad<<<(rows+ROWS_PER_CTA-1)/ROWS_PER_CTA, dim3(32, ROWS_PER_CTA), 0, stream>>>(a);
    k_pmeInterpBrick<<<(a.numAtoms+31)/32, 32, 0, stream>>>(a, map);
}
