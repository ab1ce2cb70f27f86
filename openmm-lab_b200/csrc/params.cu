// Device-side computeParameters: particle and exception parameters with their global-parameter offsets, and the
// per-subset Ewald self energies, in one launch with no host synchronisation.  Runs only when the base parameters
// were (re)loaded or an offset parameter changed value -- the alchemical protocols that move an offset parameter
// every step pay one small kernel instead of an O(N) host loop and four synchronous uploads.
//
// What it computes (FP64 throughout, as the Reference platform):
//   particles   q = q0 + sum p*dq, sigma = sigma0 + sum p*dsigma, eps = eps0 + sum p*deps
//               -> (sigma/2, 2 sqrt(eps), q)                      ReferenceOpenMMLabKernels.cpp:344-361
//   exceptions  the same for (chargeProd, sigma, eps) -> (sigma, 4 eps, chargeProd)          :365-384
//   self energy E[diag(s)][Coul] -= k q^2 alpha/sqrt(pi);  LJPME: E[diag(s)][vdW] += alpha_d^6 64 (sigma/2)^6 (2 sqrt eps)^2 / 12
//                                                                 ReferenceSlicedLJCoulombIxn.cpp:198-206
// (GPU behavioural cross-reference: platforms/common/src/kernels/nonbondedParameters.cc:4-78; the reference's second
// kernel, computeExclusionParameters :83-108, has no counterpart here because k_bonded reads the particle parameters
// directly.)  Offsets arrive as a CSR over the particle / exception index, one entry per (parameter, index) -- the
// reference keeps them in a map with that key, so a repeated key keeps its last scale; the host builds the CSR that way.
//
// The self-energy sums are deterministic: a fixed grid, a fixed per-thread stride, a tree inside the block and a final
// pass over the per-block partial sums in block order by the last block to finish.
#include "kernels.cuh"

#define PARAM_THREADS 256

__global__ void __launch_bounds__(PARAM_THREADS) k_computeParameters(ParamArgs a) {
    __shared__ double sSelf[2*SNB_MAX_SUBSETS][PARAM_THREADS/32];
    __shared__ bool sLast;
    const int stride = gridDim.x*blockDim.x, first = blockIdx.x*blockDim.x+threadIdx.x;
    double selfC[SNB_MAX_SUBSETS], selfV[SNB_MAX_SUBSETS];
#pragma unroll
    for (int s = 0; s < SNB_MAX_SUBSETS; s++)
        selfC[s] = selfV[s] = 0.0;
    for (int i = first; i < a.numAtoms; i += stride) {
        double q = a.baseParticle[3*(size_t) i], sigma = a.baseParticle[3*(size_t) i+1], eps = a.baseParticle[3*(size_t) i+2];
        if (a.particleOffsetStart)
            for (int k = a.particleOffsetStart[i]; k < a.particleOffsetStart[i+1]; k++) {
                const double4 o = a.particleOffsets[k];
                const double value = a.globals[(int) o.w];
                q += value*o.x; sigma += value*o.y; eps += value*o.z;
            }
        const double sig = 0.5*sigma, e2 = 2.0*sqrt(eps);
        a.paramsD[3*(size_t) i] = sig; a.paramsD[3*(size_t) i+1] = e2; a.paramsD[3*(size_t) i+2] = q;
        const int sub = a.subset[i];
        a.params4[i] = make_float4((float) q, (float) sig, (float) e2, __int_as_float(sub));
        const double sig2 = sig*sig, sig6 = sig2*sig2*sig2;
#pragma unroll
        for (int s = 0; s < SNB_MAX_SUBSETS; s++)
            if (s == sub) {
                selfC[s] -= a.selfScaleC*q*q;
                selfV[s] += a.selfScaleV*sig6*e2*e2;
            }
    }
    for (int e = first; e < a.numExceptions; e += stride) {
        double qq = a.baseException[3*(size_t) e], sigma = a.baseException[3*(size_t) e+1], eps = a.baseException[3*(size_t) e+2];
        if (a.exceptionOffsetStart)
            for (int k = a.exceptionOffsetStart[e]; k < a.exceptionOffsetStart[e+1]; k++) {
                const double4 o = a.exceptionOffsets[k];
                const double value = a.globals[(int) o.w];
                qq += value*o.x; sigma += value*o.y; eps += value*o.z;
            }
        a.exception14[3*(size_t) e] = sigma; a.exception14[3*(size_t) e+1] = 4.0*eps; a.exception14[3*(size_t) e+2] = qq;
    }
    // ---- self energies: warp tree, block tree, per-block partial sums, last block adds them in block order
    const int lane = threadIdx.x&31, warp = threadIdx.x>>5;
#pragma unroll
    for (int s = 0; s < SNB_MAX_SUBSETS; s++) {
        double c = selfC[s], v = selfV[s];
        for (int o = 16; o > 0; o >>= 1) {
            c += __shfl_down_sync(0xffffffffu, c, o);
            v += __shfl_down_sync(0xffffffffu, v, o);
        }
        if (lane == 0) { sSelf[2*s][warp] = c; sSelf[2*s+1][warp] = v; }
    }
    __syncthreads();
    if (threadIdx.x < 2*SNB_MAX_SUBSETS) {
        double sum = 0.0;
        for (int w = 0; w < PARAM_THREADS/32; w++)
            sum += sSelf[threadIdx.x][w];
        a.selfPartial[(size_t) blockIdx.x*2*SNB_MAX_SUBSETS+threadIdx.x] = sum;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        sLast = atomicAdd(a.ticket, 1u) == gridDim.x-1;
    __syncthreads();
    if (sLast) {
        __threadfence();
        if (threadIdx.x < 2*SNB_MAX_SUBSETS) {
            double sum = 0.0;
            for (unsigned b = 0; b < gridDim.x; b++)
                sum += ((volatile double*) a.selfPartial)[(size_t) b*2*SNB_MAX_SUBSETS+threadIdx.x];
            a.selfOut[threadIdx.x] = sum;       // [subset][Coulomb, vdW]
        }
        if (threadIdx.x == 0)
            *a.ticket = 0u;                     // ready for the next launch
    }
}

int paramGridSize(int numAtoms, int numExceptions) {
    const int n = numAtoms > numExceptions ? numAtoms : numExceptions;
    int blocks = (n+PARAM_THREADS-1)/PARAM_THREADS;
    if (blocks < 1) blocks = 1;
    return blocks > SNB_PARAM_MAX_BLOCKS ? SNB_PARAM_MAX_BLOCKS : blocks;
}

void launchComputeParameters(const ParamArgs& a, cudaStream_t stream) {
    k_computeParameters<<<paramGridSize(a.numAtoms, a.numExceptions), PARAM_THREADS, 0, stream>>>(a);
}
