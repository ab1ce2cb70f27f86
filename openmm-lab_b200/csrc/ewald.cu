// Classic Ewald summation, reciprocal space: the O(N K) sum over the half space of
// reciprocal vectors with per-subset structure factors, so that the energy decomposes into subset-pair slices and
// every force term carries the slice's Coulomb lambda.  FP64 throughout: the method is for small systems (PME is
// the path for everything else), and it is here for the reference's six-method matrix, not for speed.
//
// Follows platforms/reference/src/ReferenceSlicedLJCoulombIxn.cpp:240-342:
//   S_s(k) = sum_{n in subset s} q_n exp(i k.r_n),  a_k = exp(-k^2/(4 alpha^2))/k^2,  c = k_e 4 pi/V
//   E[slice(j,j)] += c a_k |S_j|^2,   E[slice(j,i<j)] += 2 c a_k Re(S_i conj S_j)                       (:331-335)
//   F_n += 2 c a_k k sum_j lambda_C[slice(s_n,j)] (Re S_j Im t_n - Im S_j Re t_n),  t_n = q_n exp(i k.r_n)   (:320-329)
// over k = 2 pi (rx/Lx, ry/Ly, rz/Lz): rx = 0..kmax_x-1; ry, rz on both sides except the half planes the loop skips
// (:279-296).  The reference builds exp(i k.r) by recurrence; here every phase is evaluated directly (sincos).
// The self term -k_e q^2 alpha/sqrt(pi) is added on the host (api.cu finishEnergies), the exclusion terms by k_bonded.
#include "kernels.cuh"

#define EWALD_THREADS 128
#define EWALD_KSPLIT 8          /* the force kernel splits the k vectors over blockIdx.y */

namespace {

__device__ __forceinline__ double3 kVector(const int3 r, const BoxD& b) {
    const double twoPi = 6.283185307179586;
    return make_double3(r.x*(twoPi/b.ax), r.y*(twoPi/b.by), r.z*(twoPi/b.cz));
}

// one CTA per k vector: structure factors of every subset, and the k vector's energy terms
__global__ void __launch_bounds__(EWALD_THREADS) k_ewaldStructure(EwaldArgs a) {
    __shared__ double sSum[EWALD_THREADS/32][2*SNB_MAX_SUBSETS];
    const BoxD& b = a.sp->exact.boxd;
    const double3 k = kVector(a.kvec[blockIdx.x], b);
    double acc[2*SNB_MAX_SUBSETS];
#pragma unroll
    for (int s = 0; s < 2*SNB_MAX_SUBSETS; s++)
        acc[s] = 0.0;
    for (int n = threadIdx.x; n < a.numAtoms; n += EWALD_THREADS) {
        double sn, cs;
        sincos(k.x*a.posd[3*n]+k.y*a.posd[3*n+1]+k.z*a.posd[3*n+2], &sn, &cs);
        const double q = a.paramsD[3*n+2];
        const int sub = a.subset[n];
#pragma unroll
        for (int s = 0; s < SNB_MAX_SUBSETS; s++) {
            acc[2*s] += sub == s ? q*cs : 0.0;
            acc[2*s+1] += sub == s ? q*sn : 0.0;
        }
    }
    const int lane = threadIdx.x&31, warp = threadIdx.x>>5;
#pragma unroll
    for (int s = 0; s < 2*SNB_MAX_SUBSETS; s++) {
        double v = acc[s];
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0)
            sSum[warp][s] = v;
    }
    __syncthreads();
    if (threadIdx.x < 2*a.numSubsets) {
        double v = 0.0;
        for (int w = 0; w < EWALD_THREADS/32; w++)
            v += sSum[w][threadIdx.x];
        sSum[0][threadIdx.x] = v;
        ((double*) a.structure)[(size_t) blockIdx.x*2*a.numSubsets+threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && a.energyBuf) {
        const double k2 = k.x*k.x+k.y*k.y+k.z*k.z;
        const double ak = exp(-k2/(4.0*a.alpha*a.alpha))/k2;
        const double coeff = SNB_ONE_4PI_EPS0*4.0*M_PI/(b.ax*b.by*b.cz)*ak;
        double* table = a.energyBuf+(size_t) (blockIdx.x%32)*2*SNB_MAX_SLICES;
        for (int j = 0; j < a.numSubsets; j++) {
            const double cj = sSum[0][2*j], sj = sSum[0][2*j+1];
            for (int i = 0; i < j; i++)
                atomicAdd(&table[2*(j*(j+1)/2+i)], 2.0*coeff*(sSum[0][2*i]*cj+sSum[0][2*i+1]*sj));
            atomicAdd(&table[2*(j*(j+3)/2)], coeff*(cj*cj+sj*sj));
        }
    }
}

// one thread per atom and share of the k vectors: forces, in 2^32 fixed point (sums are order independent)
__global__ void __launch_bounds__(EWALD_THREADS) k_ewaldForces(EwaldArgs a) {
    const int n = blockIdx.x*EWALD_THREADS+threadIdx.x;
    if (n >= a.numAtoms)
        return;
    const BoxD& b = a.sp->exact.boxd;
    const double x = a.posd[3*n], y = a.posd[3*n+1], z = a.posd[3*n+2], q = a.paramsD[3*n+2];
    const int sub = a.subset[n];
    double lam[SNB_MAX_SUBSETS];
#pragma unroll
    for (int j = 0; j < SNB_MAX_SUBSETS; j++)
        lam[j] = j < a.numSubsets ? a.sp->lambdaC[sliceOf(sub, j)] : 0.0;
    const double coeff = 2.0*SNB_ONE_4PI_EPS0*4.0*M_PI/(b.ax*b.by*b.cz);
    const double factor = -1.0/(4.0*a.alpha*a.alpha);
    const int per = (a.numK+EWALD_KSPLIT-1)/EWALD_KSPLIT;
    const int kBegin = min(a.numK, (int) blockIdx.y*per), kEnd = min(a.numK, kBegin+per);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int kk = kBegin; kk < kEnd; kk++) {
        const double3 k = kVector(a.kvec[kk], b);
        double sn, cs;
        sincos(k.x*x+k.y*y+k.z*z, &sn, &cs);
        const double tr = q*cs, ti = q*sn;
        const double k2 = k.x*k.x+k.y*k.y+k.z*k.z;
        const double ak = exp(k2*factor)/k2;
        const double* S = (const double*) a.structure+(size_t) kk*2*a.numSubsets;
        double f = 0.0;
#pragma unroll
        for (int j = 0; j < SNB_MAX_SUBSETS; j++)
            if (j < a.numSubsets)
                f += lam[j]*(S[2*j]*ti-S[2*j+1]*tr);
        f *= coeff*ak;
        fx += f*k.x; fy += f*k.y; fz += f*k.z;
    }
    addFixed(&a.forceOrig[n], toFixedD(fx));
    addFixed(&a.forceOrig[a.paddedAtoms+n], toFixedD(fy));
    addFixed(&a.forceOrig[2*a.paddedAtoms+n], toFixedD(fz));
}

}  // namespace

void launchEwaldReciprocal(const EwaldArgs& a, bool includeForces, cudaStream_t stream) {
    if (a.numK <= 0 || a.numAtoms <= 0)
        return;
    k_ewaldStructure<<<a.numK, EWALD_THREADS, 0, stream>>>(a);
    if (includeForces)
        k_ewaldForces<<<dim3((a.numAtoms+EWALD_THREADS-1)/EWALD_THREADS, EWALD_KSPLIT), EWALD_THREADS, 0, stream>>>(a);
}
