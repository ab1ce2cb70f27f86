// The CUDA programming guide's own TMA examples, to tell whether bulk-async copies work on this box at all:
// (1) 1-D bulk copy global->shared with an mbarrier (UBLKCP), (2) 2-D tensor tile load (UTMALDG), both through libcu++.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__global__ void k_bulk1d(int* data, int* out) {
    __shared__ alignas(16) int smem[1024];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_global_to_shared(smem, data, sizeof(smem), bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    out[threadIdx.x] = smem[threadIdx.x];
}

__global__ void k_tensor2d(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
    __shared__ alignas(128) int smem[32][32];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    out[threadIdx.x] = smem[threadIdx.x/32][threadIdx.x%32];
}

__global__ void k_tensor3d(const __grid_constant__ CUtensorMap tensor_map, int c0, int c1, int c2, float* out) {
    __shared__ alignas(128) float smem[16*16*16];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_3d_global_to_shared(smem, &tensor_map, c0, c1, c2, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int k = threadIdx.x; k < 4096; k += blockDim.x) out[k] = smem[k];
}

static int test3d(int which) {
    const int nzp = 24, py = 22, px = 44;
    float *grid, *out;
    cudaMalloc(&grid, (size_t) nzp*py*px*4); cudaMalloc(&out, 4096*4);
    float* h = (float*) malloc((size_t) nzp*py*px*4);
    for (int i = 0; i < nzp*py*px; i++) h[i] = (float) i;
    cudaMemcpy(grid, h, (size_t) nzp*py*px*4, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult st;
    if (which == 2) cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &st);
    else cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st);
    CUtensorMap map{};
    cuuint64_t size[3] = {(cuuint64_t) nzp, (cuuint64_t) py, (cuuint64_t) px}; cuuint64_t stride[2] = {(cuuint64_t) nzp*4, (cuuint64_t) nzp*py*4};
    cuuint32_t box[3] = {16, 16, 16}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiled) fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, grid, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode3d %d fn %p\n", (int) r, fn);
    k_tensor3d<<<1, 128>>>(map, 2, 3, 1, out);
    printf("tensor 3d (variant %d): %s\n", which, cudaGetErrorString(cudaDeviceSynchronize()));
    float ho[4]; cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
    printf("out[0..1] = %g %g (expect %g..)\n", ho[0], ho[1], (float) ((1*py+3)*nzp+2));
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && atoi(argv[1]) >= 2) return test3d(atoi(argv[1]));
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    int *data, *out;
    cudaMalloc(&data, 256*256*4); cudaMalloc(&out, 1024*4);
    int* h = (int*) malloc(256*256*4);
    for (int i = 0; i < 256*256; i++) h[i] = i;
    cudaMemcpy(data, h, 256*256*4, cudaMemcpyHostToDevice);
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    printf("device %s cc %d.%d\n", p.name, p.major, p.minor);
    if (which == 0) {
        k_bulk1d<<<1, 128>>>(data, out);
        printf("bulk 1d: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    } else {
        typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr; cudaDriverEntryPointQueryResult st;
        cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &st);
        CUtensorMap map{};
        cuuint64_t size[2] = {256, 256}; cuuint64_t stride[1] = {256*4}; cuuint32_t box[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
        CUresult r = ((EncodeTiled) fn)(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, data, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d\n", (int) r);
        k_tensor2d<<<1, 128>>>(map, 32, 64, out);
        printf("tensor 2d: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        int ho[128]; cudaMemcpy(ho, out, 512, cudaMemcpyDeviceToHost);
        printf("out[0..3] = %d %d %d %d (expect %d..)\n", ho[0], ho[1], ho[2], ho[3], 64*256+32);
    }
    return 0;
}
