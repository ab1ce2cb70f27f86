// Which TMA tile shapes load on this box?  usage: tma_shape_test rank isFloat b0 b1 b2 d0 d1 d2
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__global__ void k(const __grid_constant__ CUtensorMap map, int rank, int bytes, int c0, int* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (rank == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(smem, &map, c0, 2, bar);
        else cde::cp_async_bulk_tensor_3d_global_to_shared(smem, &map, c0, 2, 3, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    if (threadIdx.x < 4) out[threadIdx.x] = ((int*) smem)[threadIdx.x];
}

int main(int argc, char** argv) {
    const int rank = atoi(argv[1]), isFloat = atoi(argv[2]);
    cuuint32_t box[3] = {(cuuint32_t) atoi(argv[3]), (cuuint32_t) atoi(argv[4]), (cuuint32_t) atoi(argv[5])};
    cuuint64_t dims[3] = {(cuuint64_t) atoi(argv[6]), (cuuint64_t) atoi(argv[7]), (cuuint64_t) atoi(argv[8])};
    size_t total = dims[0]*dims[1]*dims[2];
    int* data; int* out;
    cudaMalloc(&data, total*4); cudaMalloc(&out, 16);
    int* h = (int*) malloc(total*4);
    for (size_t i = 0; i < total; i++) { if (isFloat) ((float*) h)[i] = (float) i; else h[i] = (int) i; }
    cudaMemcpy(data, h, total*4, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult st;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st);
    CUtensorMap map{};
    cuuint64_t stride[2] = {dims[0]*4, dims[0]*dims[1]*4}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiled) fn)(&map, isFloat ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_INT32, rank, data, dims, stride, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int bytes = box[0]*box[1]*(rank == 3 ? box[2] : 1)*4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    const int c0 = argc > 9 ? atoi(argv[9]) : 1;
    k<<<1, 128, bytes>>>(map, rank, bytes, c0, out);
    cudaError_t e = cudaDeviceSynchronize();
    int ho[4] = {0, 0, 0, 0}; cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
    size_t first = rank == 2 ? 2*dims[0]+c0 : (3*dims[1]+2)*dims[0]+c0;
    printf("c0 %d rank %d %s box %u %u %u dims %llu %llu %llu: encode %d, %s, first %d (expect %zu)\n", c0, rank, isFloat ? "float" : "int", box[0], box[1], box[2],
           (unsigned long long) dims[0], (unsigned long long) dims[1], (unsigned long long) dims[2], (int) r, cudaGetErrorString(e), isFloat ? (int) ((float*) ho)[0] : ho[0], first);
    return 0;
}
