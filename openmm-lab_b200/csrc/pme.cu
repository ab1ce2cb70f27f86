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

#include <algorithm>

#define ORDER SNB_PME_ORDER
#define PME_FIXED_SCALE 1099511627776.0      /* 2^40: deterministic (order-independent) spreading */

template <typename R> __device__ __forceinline__ void bspline5(R w, R* th, R* dth) {
    // Essmann recursion for order 5 (ReferencePME.cpp:264-317)
    th[4] = 0;
    th[1] = w;
    th[0] = 1-w;
#pragma unroll
    for (int k = 3; k < ORDER; k++) {
        const R div = (R) 1/(R) (k-1);
        th[k-1] = div*w*th[k-2];
#pragma unroll
        for (int l = 1; l < k-1; l++)
            th[k-l-1] = div*((w+l)*th[k-l-2]+(k-l-w)*th[k-l-1]);
        th[0] = div*(1-w)*th[0];
    }
    dth[0] = -th[0];
#pragma unroll
    for (int k = 1; k < ORDER; k++)
        dth[k] = th[k-1]-th[k];
    const R div = (R) 1/(R) (ORDER-1);
    th[ORDER-1] = div*w*th[ORDER-2];
#pragma unroll
    for (int l = 1; l < ORDER-1; l++)
        th[ORDER-l-1] = div*((w+l)*th[ORDER-l-2]+(ORDER-l-w)*th[ORDER-l-1]);
    th[0] = div*(1-w)*th[0];
}

// fractional grid coordinate in double (positions may be far from the origin)
template <typename R> __device__ __forceinline__ void gridCoord(const PmeArgs& a, int atom, int idx[3], R frac[3]) {
    const double x = a.posd[3*atom], y = a.posd[3*atom+1], z = a.posd[3*atom+2];
    double t[3];
    t[0] = x*a.sp->exact.boxd.recip[0][0]+y*a.sp->exact.boxd.recip[1][0]+z*a.sp->exact.boxd.recip[2][0];
    t[1] = y*a.sp->exact.boxd.recip[1][1]+z*a.sp->exact.boxd.recip[2][1];
    t[2] = z*a.sp->exact.boxd.recip[2][2];
    const int n[3] = {a.nx, a.ny, a.nz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double u = (t[d]-floor(t[d]))*n[d];
        int ti = (int) u;
        frac[d] = (R) (u-ti);
        idx[d] = ti%n[d];
    }
}

template <typename R> __device__ __forceinline__ R pmeCharge(const PmeArgs& a, int atom) {
    if (sizeof(R) == 8) {
        const double sg = a.paramsD[3*atom], ep = a.paramsD[3*atom+1];
        return (R) (a.term == 0 ? a.paramsD[3*atom+2] : 8.0*sg*sg*sg*ep);
    }
    const float4 prm = a.params4[atom];
    return (R) (a.term == 0 ? prm.x : 8.0f*prm.y*prm.y*prm.y*prm.z);
}

// ---- spreading: one warp per atom (sorted order: neighbouring warps hit neighbouring cells);
// lane -> (ix, iy) for the 25 xy offsets of the 5x5x5 stencil, 5 consecutive z points each.
template <typename R, bool FIXED>
__global__ void __launch_bounds__(256) k_pmeSpread(PmeArgs a) {
    const int k = (blockIdx.x*blockDim.x+threadIdx.x)>>5, lane = threadIdx.x&31;
    if (k >= a.numAtoms || lane >= ORDER*ORDER)
        return;
    const int atom = a.sortedAtom ? a.sortedAtom[k] : k;
    if (atom < 0)
        return;
    const R q = pmeCharge<R>(a, atom);
    if (q == 0)
        return;
    int idx[3];
    R frac[3], tx[ORDER], ty[ORDER], tz[ORDER], dt[ORDER];
    gridCoord<R>(a, atom, idx, frac);
    bspline5<R>(frac[0], tx, dt);
    bspline5<R>(frac[1], ty, dt);
    bspline5<R>(frac[2], tz, dt);
    const int subset = __float_as_int(a.params4[atom].w);
    const int ix = lane/ORDER, iy = lane-ix*ORDER;
    int x = idx[0]+ix, y = idx[1]+iy;
    x -= x >= a.nx ? a.nx : 0;
    y -= y >= a.ny ? a.ny : 0;
    R wx = tx[0], wy = ty[0];
#pragma unroll
    for (int m = 1; m < ORDER; m++) {
        wx = ix == m ? tx[m] : wx;
        wy = iy == m ? ty[m] : wy;
    }
    const R qxy = q*wx*wy;
    const size_t row = (((size_t) subset*a.nx+x)*a.ny+y)*a.nz;
#pragma unroll
    for (int iz = 0; iz < ORDER; iz++) {
        int z = idx[2]+iz;
        z -= z >= a.nz ? a.nz : 0;
        if (FIXED)
            atomicAdd((unsigned long long*) a.gridFixed+row+z, (unsigned long long) __double2ll_rn((double) (qxy*tz[iz])*PME_FIXED_SCALE));
        else
            atomicAdd((R*) a.grid+row+z, qxy*tz[iz]);
    }
}

template <typename R>
__global__ void k_pmeFinishSpread(PmeArgs a, size_t total) {
    for (size_t i = blockIdx.x*(size_t) blockDim.x+threadIdx.x; i < total; i += (size_t) gridDim.x*blockDim.x)
        ((R*) a.grid)[i] = (R) ((double) a.gridFixed[i]*(1.0/PME_FIXED_SCALE));
}

// ---- reciprocal-space kernel table: eterm[kx][ky][kz<=nz/2], rebuilt only when the box changes
template <typename R>
__global__ void k_pmeEterm(PmeArgs a, R* eterm) {
    const int nzc = a.nz/2+1;
    const size_t total = (size_t) a.nx*a.ny*nzc;
    const double volume = a.sp->exact.boxd.ax*a.sp->exact.boxd.by*a.sp->exact.boxd.cz;
    for (size_t idx = blockIdx.x*(size_t) blockDim.x+threadIdx.x; idx < total; idx += (size_t) gridDim.x*blockDim.x) {
        const int kz = idx%nzc, ky = (idx/nzc)%a.ny, kx = idx/((size_t) nzc*a.ny);
        const double mx = kx < (a.nx+1)/2 ? kx : kx-a.nx;
        const double my = ky < (a.ny+1)/2 ? ky : ky-a.ny;
        const double mz = kz < (a.nz+1)/2 ? kz : kz-a.nz;
        const double mhx = mx*a.sp->exact.boxd.recip[0][0];
        const double mhy = mx*a.sp->exact.boxd.recip[1][0]+my*a.sp->exact.boxd.recip[1][1];
        const double mhz = mx*a.sp->exact.boxd.recip[2][0]+my*a.sp->exact.boxd.recip[2][1]+mz*a.sp->exact.boxd.recip[2][2];
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
        eterm[idx] = (R) e;
    }
}

template <typename R> struct Cplx;
template <> struct Cplx<float> { typedef float2 type; };
template <> struct Cplx<double> { typedef double2 type; };

// ---- convolution + per-slice energies on the half-complex grid
template <typename R>
__global__ void __launch_bounds__(256) k_pmeConvolution(PmeArgs a) {
    typedef typename Cplx<R>::type C2;
    __shared__ double sE[SNB_MAX_SLICES];
    for (int k = threadIdx.x; k < SNB_MAX_SLICES; k += blockDim.x)
        sE[k] = 0.0;
    __syncthreads();
    const int nzc = a.nz/2+1, S = a.numSubsets;
    const size_t total = (size_t) a.nx*a.ny*nzc;
    C2* cgrid = (C2*) a.cgrid;
    const R* eterm = (const R*) a.eterm;
    double acc[SNB_MAX_SLICES];
    const int L = S*(S+1)/2;
    for (int k = 0; k < L; k++)
        acc[k] = 0.0;
    for (size_t idx = blockIdx.x*(size_t) blockDim.x+threadIdx.x; idx < total; idx += (size_t) gridDim.x*blockDim.x) {
        const int kz = idx%nzc;
        const R e = eterm[idx];
        // Hermitian fold: planes 0 < kz < nz/2 stand for two grid points of the full sum
        const R wgt = (kz == 0 || (2*kz == a.nz)) ? 1 : 2;
        C2 v[SNB_MAX_SUBSETS];
        for (int j = 0; j < S; j++) {
            v[j] = cgrid[(size_t) j*total+idx];
            for (int i = 0; i < j; i++)
                acc[j*(j+1)/2+i] += (double) (wgt*e*(v[j].x*v[i].x+v[j].y*v[i].y));
            acc[j*(j+3)/2] += (double) ((R) 0.5*wgt*e*(v[j].x*v[j].x+v[j].y*v[j].y));
            C2 o;
            o.x = v[j].x*e; o.y = v[j].y*e;
            cgrid[(size_t) j*total+idx] = o;
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

// ---- interpolation: one warp per atom, lane -> (ix, iy), 5 z points x all subsets' potentials
template <typename R>
__global__ void __launch_bounds__(256) k_pmeInterpolate(PmeArgs a) {
    const int k = (blockIdx.x*blockDim.x+threadIdx.x)>>5, lane = threadIdx.x&31;
    if (k >= a.numAtoms)
        return;
    const int atom = a.sortedAtom ? a.sortedAtom[k] : k;
    if (atom < 0)
        return;
    const R q = pmeCharge<R>(a, atom);
    if (q == 0)
        return;
    double fx = 0, fy = 0, fz = 0;
    if (lane < ORDER*ORDER) {
        int idx[3];
        R frac[3], tx[ORDER], ty[ORDER], tz[ORDER], dtx[ORDER], dty[ORDER], dtz[ORDER];
        gridCoord<R>(a, atom, idx, frac);
        bspline5<R>(frac[0], tx, dtx);
        bspline5<R>(frac[1], ty, dty);
        bspline5<R>(frac[2], tz, dtz);
        const int si = __float_as_int(a.params4[atom].w);
        const size_t G = (size_t) a.nx*a.ny*a.nz;
        const int ix = lane/ORDER, iy = lane-ix*ORDER;
        int x = idx[0]+ix, y = idx[1]+iy;
        x -= x >= a.nx ? a.nx : 0;
        y -= y >= a.ny ? a.ny : 0;
        R wx = tx[0], wy = ty[0], dwx = dtx[0], dwy = dty[0];
#pragma unroll
        for (int m = 1; m < ORDER; m++) {
            wx = ix == m ? tx[m] : wx; dwx = ix == m ? dtx[m] : dwx;
            wy = iy == m ? ty[m] : wy; dwy = iy == m ? dty[m] : dwy;
        }
        const R* row = (const R*) a.grid+((size_t) x*a.ny+y)*a.nz;
        R s0 = 0, s1 = 0;           // sum_z theta_z * phi, sum_z dtheta_z * phi
        for (int sj = 0; sj < a.numSubsets; sj++) {
            const R lam = (R) (a.term == 0 ? a.sp->lambdaC : a.sp->lambdaV)[sliceOf(si, sj)];
#pragma unroll
            for (int iz = 0; iz < ORDER; iz++) {
                int z = idx[2]+iz;
                z -= z >= a.nz ? a.nz : 0;
                const R v = lam*row[sj*G+z];
                s0 += tz[iz]*v;
                s1 += dtz[iz]*v;
            }
        }
        fx = dwx*wy*s0;
        fy = wx*dwy*s0;
        fz = wx*wy*s1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (lane == 0) {
        // ReferencePME.cpp:698-700
        const double qd = q;
        const double gx = fx*a.nx, gy = fy*a.ny, gz = fz*a.nz;
        const double Fx = -qd*(gx*a.sp->exact.boxd.recip[0][0]);
        const double Fy = -qd*(gx*a.sp->exact.boxd.recip[1][0]+gy*a.sp->exact.boxd.recip[1][1]);
        const double Fz = -qd*(gx*a.sp->exact.boxd.recip[2][0]+gy*a.sp->exact.boxd.recip[2][1]+gz*a.sp->exact.boxd.recip[2][2]);
        addFixed(&a.forceOrig[atom], toFixedD(Fx));
        addFixed(&a.forceOrig[a.paddedAtoms+atom], toFixedD(Fy));
        addFixed(&a.forceOrig[2*a.paddedAtoms+atom], toFixedD(Fz));
    }
}

void launchPmeEterm(const PmeArgs& a, cudaStream_t stream) {
    if (a.doublePrecision) k_pmeEterm<double><<<592, 256, 0, stream>>>(a, (double*) a.eterm);
    else k_pmeEterm<float><<<592, 256, 0, stream>>>(a, (float*) a.eterm);
}
void launchPmeSpread(const PmeArgs& a, cudaStream_t stream) {
    const size_t total = (size_t) a.numSubsets*a.nx*a.ny*a.nz;
    const int blocks = (int) (((size_t) a.numAtoms*32+255)/256);
    if (a.gridFixed) {
        cudaMemsetAsync(a.gridFixed, 0, sizeof(long long)*total, stream);
        int fb = (int) std::min<size_t>((total+255)/256, 2368);
        if (a.doublePrecision) {
            k_pmeSpread<double, true><<<blocks, 256, 0, stream>>>(a);
            k_pmeFinishSpread<double><<<fb, 256, 0, stream>>>(a, total);
        }
        else {
            k_pmeSpread<float, true><<<blocks, 256, 0, stream>>>(a);
            k_pmeFinishSpread<float><<<fb, 256, 0, stream>>>(a, total);
        }
    }
    else {
        cudaMemsetAsync(a.grid, 0, (a.doublePrecision ? 8 : 4)*total, stream);
        if (a.doublePrecision) k_pmeSpread<double, false><<<blocks, 256, 0, stream>>>(a);
        else k_pmeSpread<float, false><<<blocks, 256, 0, stream>>>(a);
    }
}
void launchPmeConvolution(const PmeArgs& a, cudaStream_t stream) {
    size_t total = (size_t) a.nx*a.ny*(a.nz/2+1);
    int blocks = (int) std::min<size_t>((total+255)/256, 1184);
    if (a.doublePrecision) k_pmeConvolution<double><<<blocks, 256, 0, stream>>>(a);
    else k_pmeConvolution<float><<<blocks, 256, 0, stream>>>(a);
}
void launchPmeInterpolate(const PmeArgs& a, cudaStream_t stream) {
    const int blocks = (int) (((size_t) a.numAtoms*32+255)/256);
    if (a.doublePrecision) k_pmeInterpolate<double><<<blocks, 256, 0, stream>>>(a);
    else k_pmeInterpolate<float><<<blocks, 256, 0, stream>>>(a);
}
