// Stand-alone check of the TMA brick transfers used by pme_brick.cu: one 16x16x16 float box loaded from / reduced into a
// padded 3-D grid.  Prints the CUDA error after every kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_brick_test tma_brick_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <climits>
#include <cstdlib>

#define BRICK 16
#define BRICK_FLOATS (BRICK*BRICK*BRICK)

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

__global__ void k_redux(int* out) {
    const int lane = threadIdx.x;
    const int lo = __reduce_min_sync(0xffffffffu, lane == 5 ? -7 : lane), hi = __reduce_max_sync(0xffffffffu, lane);
    if (lane == 0) { out[0] = lo; out[1] = hi; }
}

__global__ void __launch_bounds__(32) k_loadPtr(const CUtensorMap* tmapPtr, int cz, int cy, int cx, float* out) {
    __shared__ __align__(128) float sBrick[BRICK_FLOATS];
    __shared__ __align__(8) unsigned long long sBar;
    const int lane = threadIdx.x;
    const unsigned bar = smemAddr(&sBar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "n"(BRICK_FLOATS*4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smemAddr(sBrick)), "l"((unsigned long long) tmapPtr), "r"(bar), "r"(cz), "r"(cy), "r"(cx) : "memory");
    }
    unsigned done = 0;
    for (int spin = 0; !done && spin < (1<<20); spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    if (lane == 0) out[BRICK_FLOATS] = done ? 1.f : 0.f;
    for (int k = lane; k < BRICK_FLOATS; k += 32)
        out[k] = sBrick[k];
}

__global__ void __launch_bounds__(32) k_load(const __grid_constant__ CUtensorMap tmap, int cz, int cy, int cx, float* out) {
    __shared__ __align__(128) float sBrick[BRICK_FLOATS];
    __shared__ __align__(8) unsigned long long sBar;
    const int lane = threadIdx.x;
    const unsigned bar = smemAddr(&sBar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "n"(BRICK_FLOATS*4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smemAddr(sBrick)), "l"((unsigned long long) &tmap), "r"(bar), "r"(cz), "r"(cy), "r"(cx) : "memory");
    }
    unsigned done = 0;
    for (int spin = 0; !done && spin < (1<<20); spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    if (lane == 0) out[BRICK_FLOATS] = done ? 1.f : 0.f;
    for (int k = lane; k < BRICK_FLOATS; k += 32)
        out[k] = sBrick[k];
}

__global__ void __launch_bounds__(32) k_reduce(const __grid_constant__ CUtensorMap tmap, int cz, int cy, int cx) {
    __shared__ __align__(128) float sBrick[BRICK_FLOATS];
    const int lane = threadIdx.x;
    for (int k = lane; k < BRICK_FLOATS; k += 32)
        sBrick[k] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
                     :: "l"((unsigned long long) &tmap), "r"(smemAddr(sBrick)), "r"(cz), "r"(cy), "r"(cx) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

static void report(const char* what) {
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-28s %s\n", what, cudaGetErrorString(e));
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    printf("variant %d\n", variant);
    const int nx = 18, ny = 18, nz = 18, S = 2, nzp = (variant&4) ? 32 : (nz+4+3)&~3, py = ny+4, px = nx+4;
    const size_t total = (size_t) S*px*py*nzp;
    std::vector<float> host(total);
    for (size_t i = 0; i < total; i++) host[i] = (float) (i%1000)*0.001f;
    float *grid, *out;
    cudaMalloc(&grid, total*4);
    cudaMalloc(&out, (BRICK_FLOATS+1)*4);
    cudaMemcpy(grid, host.data(), total*4, cudaMemcpyHostToDevice);
    int* iout; cudaMalloc(&iout, 8);
    k_redux<<<1, 32>>>(iout);
    report("redux");
    int ih[2]; cudaMemcpy(ih, iout, 8, cudaMemcpyDeviceToHost); printf("  min %d max %d\n", ih[0], ih[1]);

    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult status;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status);
    printf("entry point: %s status %d fn %p\n", cudaGetErrorString(ge), (int) status, fn);
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t) nzp, (cuuint64_t) py, (cuuint64_t) S*px};
    const cuuint64_t strides[2] = {(cuuint64_t) nzp*4, (cuuint64_t) nzp*4*py};
    const cuuint32_t box[3] = {BRICK, BRICK, BRICK}, elem[3] = {1, 1, 1};
    CUresult r = ((EncodeTiled) fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, grid, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, (variant&1) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap* mapDev;
    cudaMalloc(&mapDev, sizeof(CUtensorMap));
    cudaMemcpy(mapDev, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    printf("encode: %d  (dims %d %d %d, strides %d %d)\n", (int) r, nzp, py, S*px, nzp*4, nzp*4*py);
    for (int test = 0; test < 2; test++) {
        const int cz = test ? 10 : 2, cy = test ? 9 : 3, cx = test ? px+5 : 1;
        if (variant&2) k_loadPtr<<<1, 32>>>(mapDev, cz, cy, cx, out);
        else k_load<<<1, 32>>>(map, cz, cy, cx, out);
        report("tma load");
        std::vector<float> got(BRICK_FLOATS+1);
        cudaMemcpy(got.data(), out, (BRICK_FLOATS+1)*4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int x = 0; x < BRICK; x++) for (int y = 0; y < BRICK; y++) for (int z = 0; z < BRICK; z++) {
            const int gx = cx+x, gy = cy+y, gz = cz+z;
            const float want = (gx < S*px && gy < py && gz < nzp) ? host[((size_t) gx*py+gy)*nzp+gz] : 0.f;
            if (got[(x*BRICK+y)*BRICK+z] != want) bad++;
        }
        printf("  load at (%d,%d,%d): completed %g, mismatches %d\n", cx, cy, cz, got[BRICK_FLOATS], bad);
        k_reduce<<<1, 32>>>(map, cz, cy, cx);
        report("tma reduce-add");
        std::vector<float> after(total);
        cudaMemcpy(after.data(), grid, total*4, cudaMemcpyDeviceToHost);
        bad = 0;
        for (int gx = 0; gx < S*px; gx++) for (int gy = 0; gy < py; gy++) for (int gz = 0; gz < nzp; gz++) {
            const bool in = gx >= cx && gx < cx+BRICK && gy >= cy && gy < cy+BRICK && gz >= cz && gz < cz+BRICK;
            const size_t i = ((size_t) gx*py+gy)*nzp+gz;
            if (after[i] != host[i]+(in ? 1.f : 0.f)) bad++;
            host[i] = after[i];
        }
        printf("  reduce at (%d,%d,%d): mismatches %d\n", cx, cy, cz, bad);
    }
    return 0;
}
