// Sliced PME reciprocal space (north-star subsystem 4): per-subset order-5 B-spline charge
// spreading, batched 3-D R2C/C2R FFT through cuFFT (the one library call, issued from api.cu),
// the slice-wise cross-structure-factor convolution, and force interpolation.
//
// Arithmetic follows the reference's Reference platform:
//   grid index / fraction   platforms/reference/src/ReferencePME.cpp:196-256
//   B-splines               :264-317        spreading  :320-396
//   convolution + energies  :400-496 (Coulomb), :499-595 (dispersion, includes k = 0)
//   interpolation           :598-702 (subset stride taken as nx; the reference's :682 uses nz,
//                           identical whenever nx == nz)
// (GPU behavioural cross-reference: platforms/common/src/kernels/pme.cc:27-135,179-291,296-408.)
//
// Grids are real [S][nx][ny][nz] and half-complex [S][nx][ny][nz/2+1]; the full-grid energy sums
// of the reference are folded onto the Hermitian half (weight 2 for 0 < kz < nz/2).
#include "kernels.cuh"

#define ORDER SNB_PME_ORDER

__device__ __forceinline__ void bspline5(float w, float* th, float* dth) {
    // Essmann recursion for order 5 (ReferencePME.cpp:264-317)
    th[4] = 0.f;
    th[1] = w;
    th[0] = 1.f-w;
#pragma unroll
    for (int k = 3; k < ORDER; k++) {
        const float div = 1.0f/(k-1.0f);
        th[k-1] = div*w*th[k-2];
#pragma unroll
        for (int l = 1; l < k-1; l++)
            th[k-l-1] = div*((w+l)*th[k-l-2]+(k-l-w)*th[k-l-1]);
        th[0] = div*(1.f-w)*th[0];
    }
    if (dth) {
        dth[0] = -th[0];
#pragma unroll
        for (int k = 1; k < ORDER; k++)
            dth[k] = th[k-1]-th[k];
    }
    const float div = 1.0f/(ORDER-1);
    th[ORDER-1] = div*w*th[ORDER-2];
#pragma unroll
    for (int l = 1; l < ORDER-1; l++)
        th[ORDER-l-1] = div*((w+l)*th[ORDER-l-2]+(ORDER-l-w)*th[ORDER-l-1]);
    th[0] = div*(1.f-w)*th[0];
}

// fractional grid coordinate in double (positions may be far from the origin), fraction in float
__device__ __forceinline__ void gridCoord(const PmeArgs& a, int atom, int idx[3], float frac[3]) {
    const double x = a.posd[3*atom], y = a.posd[3*atom+1], z = a.posd[3*atom+2];
    double t[3];
    t[0] = x*a.boxd.recip[0][0]+y*a.boxd.recip[1][0]+z*a.boxd.recip[2][0];
    t[1] = y*a.boxd.recip[1][1]+z*a.boxd.recip[2][1];
    t[2] = z*a.boxd.recip[2][2];
    const int n[3] = {a.nx, a.ny, a.nz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double u = (t[d]-floor(t[d]))*n[d];
        int ti = (int) u;
        frac[d] = (float) (u-ti);
        idx[d] = ti%n[d];
    }
}

__device__ __forceinline__ float pmeCharge(const PmeArgs& a, int atom) {
    const float4 prm = a.params4[atom];
    return a.term == 0 ? prm.x : 8.0f*prm.y*prm.y*prm.y*prm.z;
}

// ---- spreading: one thread per atom, in sorted order so neighbouring threads hit neighbouring cells
__global__ void __launch_bounds__(128) k_pmeSpread(PmeArgs a) {
    int k = blockIdx.x*blockDim.x+threadIdx.x;
    if (k >= a.numAtoms)
        return;
    const int atom = a.sortedAtom ? a.sortedAtom[k] : k;
    if (atom < 0)
        return;
    const float q = pmeCharge(a, atom);
    if (q == 0.f)
        return;
    int idx[3];
    float frac[3], tx[ORDER], ty[ORDER], tz[ORDER];
    gridCoord(a, atom, idx, frac);
    bspline5(frac[0], tx, nullptr);
    bspline5(frac[1], ty, nullptr);
    bspline5(frac[2], tz, nullptr);
    const int subset = __float_as_int(a.params4[atom].w);
    float* g = a.grid+(size_t) subset*a.nx*a.ny*a.nz;
    int zi[ORDER];
#pragma unroll
    for (int iz = 0; iz < ORDER; iz++) {
        int z = idx[2]+iz;
        zi[iz] = z >= a.nz ? z-a.nz : z;
    }
#pragma unroll
    for (int ix = 0; ix < ORDER; ix++) {
        int x = idx[0]+ix;
        x -= x >= a.nx ? a.nx : 0;
        const float qx = q*tx[ix];
#pragma unroll
        for (int iy = 0; iy < ORDER; iy++) {
            int y = idx[1]+iy;
            y -= y >= a.ny ? a.ny : 0;
            const float qxy = qx*ty[iy];
            float* row = g+((size_t) x*a.ny+y)*a.nz;
#pragma unroll
            for (int iz = 0; iz < ORDER; iz++)
                atomicAdd(row+zi[iz], qxy*tz[iz]);
        }
    }
}

// ---- reciprocal-space kernel table: eterm[kx][ky][kz<=nz/2], rebuilt only when the box changes
__global__ void k_pmeEterm(PmeArgs a, float* eterm) {
    const int nzc = a.nz/2+1;
    const size_t total = (size_t) a.nx*a.ny*nzc;
    const double volume = a.boxd.ax*a.boxd.by*a.boxd.cz;
    for (size_t idx = blockIdx.x*(size_t) blockDim.x+threadIdx.x; idx < total; idx += (size_t) gridDim.x*blockDim.x) {
        const int kz = idx%nzc, ky = (idx/nzc)%a.ny, kx = idx/((size_t) nzc*a.ny);
        const double mx = kx < (a.nx+1)/2 ? kx : kx-a.nx;
        const double my = ky < (a.ny+1)/2 ? ky : ky-a.ny;
        const double mz = kz < (a.nz+1)/2 ? kz : kz-a.nz;
        const double mhx = mx*a.boxd.recip[0][0];
        const double mhy = mx*a.boxd.recip[1][0]+my*a.boxd.recip[1][1];
        const double mhz = mx*a.boxd.recip[2][0]+my*a.boxd.recip[2][1]+mz*a.boxd.recip[2][2];
        const double m2 = mhx*mhx+mhy*mhy+mhz*mhz;
        const double bx = a.moduli[0][kx], by = a.moduli[1][ky], bz = a.moduli[2][kz];
        double e;
        if (a.term == 0) {
            // ReferencePME.cpp:425-427,467-471 ; k = 0 is skipped (:457-460)
            e = (kx == 0 && ky == 0 && kz == 0) ? 0.0
                : SNB_ONE_4PI_EPS0*exp(-(M_PI*M_PI/(a.alpha*a.alpha))*m2)/(m2*(M_PI*volume*bx)*by*bz);
        }
        else {
            // ReferencePME.cpp:522-570 ; k = 0 included
            const double m = sqrt(m2), b = (M_PI/a.alpha)*m;
            const double denom = (-2*M_PI*sqrt(M_PI)/(6.0*volume))/(bx*by*bz);
            e = (2.0*M_PI*M_PI*M_PI*sqrt(M_PI)*erfc(b)*m*m2+exp(-b*b)*(a.alpha*a.alpha*a.alpha-2.0*a.alpha*M_PI*M_PI*m2))*denom;
        }
        eterm[idx] = (float) e;
    }
}

// ---- convolution + per-slice energies on the half-complex grid
__global__ void __launch_bounds__(256) k_pmeConvolution(PmeArgs a) {
    __shared__ double sE[SNB_MAX_SLICES];
    for (int k = threadIdx.x; k < SNB_MAX_SLICES; k += blockDim.x)
        sE[k] = 0.0;
    __syncthreads();
    const int nzc = a.nz/2+1, S = a.numSubsets;
    const size_t total = (size_t) a.nx*a.ny*nzc;
    double acc[SNB_MAX_SLICES];
    const int L = S*(S+1)/2;
    for (int k = 0; k < L; k++)
        acc[k] = 0.0;
    for (size_t idx = blockIdx.x*(size_t) blockDim.x+threadIdx.x; idx < total; idx += (size_t) gridDim.x*blockDim.x) {
        const int kz = idx%nzc;
        const float e = a.eterm[idx];
        // Hermitian fold: planes 0 < kz < nz/2 stand for two grid points of the full sum
        const float wgt = (kz == 0 || (2*kz == a.nz)) ? 1.0f : 2.0f;
        float2 v[SNB_MAX_SUBSETS];
        for (int j = 0; j < S; j++) {
            v[j] = a.cgrid[(size_t) j*total+idx];
            for (int i = 0; i < j; i++)
                acc[j*(j+1)/2+i] += (double) (wgt*e*(v[j].x*v[i].x+v[j].y*v[i].y));
            acc[j*(j+3)/2] += (double) (0.5f*wgt*e*(v[j].x*v[j].x+v[j].y*v[j].y));
            a.cgrid[(size_t) j*total+idx] = make_float2(v[j].x*e, v[j].y*e);
        }
    }
    for (int k = 0; k < L; k++) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x&31) == 0 && v != 0.0)
            atomicAdd(&sE[k], v);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < L; k += blockDim.x)
        if (sE[k] != 0.0)
            atomicAdd(&a.energyBuf[2*k+a.term], sE[k]);
}

// ---- interpolation: one thread per atom (sorted order), all subsets' potentials
__global__ void __launch_bounds__(128) k_pmeInterpolate(PmeArgs a) {
    int k = blockIdx.x*blockDim.x+threadIdx.x;
    if (k >= a.numAtoms)
        return;
    const int atom = a.sortedAtom ? a.sortedAtom[k] : k;
    if (atom < 0)
        return;
    const float q = pmeCharge(a, atom);
    if (q == 0.f)
        return;
    int idx[3];
    float frac[3], tx[ORDER], ty[ORDER], tz[ORDER], dtx[ORDER], dty[ORDER], dtz[ORDER];
    gridCoord(a, atom, idx, frac);
    bspline5(frac[0], tx, dtx);
    bspline5(frac[1], ty, dty);
    bspline5(frac[2], tz, dtz);
    const int si = __float_as_int(a.params4[atom].w);
    const size_t G = (size_t) a.nx*a.ny*a.nz;
    float lam[SNB_MAX_SUBSETS];
    for (int sj = 0; sj < a.numSubsets; sj++)
        lam[sj] = a.lambda[sliceOf(si, sj)];
    int zi[ORDER];
#pragma unroll
    for (int iz = 0; iz < ORDER; iz++) {
        int z = idx[2]+iz;
        zi[iz] = z >= a.nz ? z-a.nz : z;
    }
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
    for (int ix = 0; ix < ORDER; ix++) {
        int x = idx[0]+ix;
        x -= x >= a.nx ? a.nx : 0;
#pragma unroll
        for (int iy = 0; iy < ORDER; iy++) {
            int y = idx[1]+iy;
            y -= y >= a.ny ? a.ny : 0;
            const float* row = a.grid+((size_t) x*a.ny+y)*a.nz;
            float s0 = 0.f, s1 = 0.f;       // sum_z theta_z * phi, sum_z dtheta_z * phi
#pragma unroll
            for (int iz = 0; iz < ORDER; iz++) {
                float v = 0.f;
                for (int sj = 0; sj < a.numSubsets; sj++)
                    v = fmaf(lam[sj], row[sj*G+zi[iz]], v);
                s0 = fmaf(tz[iz], v, s0);
                s1 = fmaf(dtz[iz], v, s1);
            }
            fx = fmaf(dtx[ix]*ty[iy], s0, fx);
            fy = fmaf(tx[ix]*dty[iy], s0, fy);
            fz = fmaf(tx[ix]*ty[iy], s1, fz);
        }
    }
    // ReferencePME.cpp:698-700
    const double qd = q;
    const double gx = fx*(double) a.nx, gy = fy*(double) a.ny, gz = fz*(double) a.nz;
    const double Fx = -qd*(gx*a.boxd.recip[0][0]);
    const double Fy = -qd*(gx*a.boxd.recip[1][0]+gy*a.boxd.recip[1][1]);
    const double Fz = -qd*(gx*a.boxd.recip[2][0]+gy*a.boxd.recip[2][1]+gz*a.boxd.recip[2][2]);
    addFixed(&a.forceOrig[atom], toFixedD(Fx));
    addFixed(&a.forceOrig[a.paddedAtoms+atom], toFixedD(Fy));
    addFixed(&a.forceOrig[2*a.paddedAtoms+atom], toFixedD(Fz));
}

void launchPmeEterm(const PmeArgs& a, float* eterm, cudaStream_t stream) {
    k_pmeEterm<<<592, 256, 0, stream>>>(a, eterm);
}
void launchPmeSpread(const PmeArgs& a, cudaStream_t stream) {
    cudaMemsetAsync(a.grid, 0, sizeof(float)*(size_t) a.numSubsets*a.nx*a.ny*a.nz, stream);
    k_pmeSpread<<<(a.numAtoms+127)/128, 128, 0, stream>>>(a);
}
void launchPmeConvolution(const PmeArgs& a, cudaStream_t stream) {
    size_t total = (size_t) a.nx*a.ny*(a.nz/2+1);
    int blocks = (int) ((total+255)/256);
    if (blocks > 1184)
        blocks = 1184;
    k_pmeConvolution<<<blocks, 256, 0, stream>>>(a);
}
void launchPmeInterpolate(const PmeArgs& a, cudaStream_t stream) {
    k_pmeInterpolate<<<(a.numAtoms+127)/128, 128, 0, stream>>>(a);
}
