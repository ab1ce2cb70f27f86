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

#include <cstdlib>

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

// ---- spreading.  A CTA takes PME_ATOMS consecutive atoms of the sorted order (neighbouring atoms hit
// neighbouring grid cells).  Phase 1, one thread per atom: grid index, order-5 B-spline weights and
// charge, once, into shared memory.  Phase 2, one warp per atom: lane -> (iy, iz) of the 5x5 yz face of
// the stencil, a serial walk over the 5 x planes, so that ONE atomic instruction of the warp covers 5
// rows x 5 consecutive z points (5-10 sectors) instead of 25 different rows.
#define PME_ATOMS 64             /* atoms per CTA */
#define PME_THREADS 256          /* 8 warps: 8 atoms each in phase 2 */
#define PME_ATOMS_PER_WARP (PME_ATOMS/(PME_THREADS/32))

template <typename R> struct PmeAtom {
    R th[3][ORDER];
    R q;
    int idx[3];
    int subset;         // -1: nothing to spread (padding, zero charge)
};

template <typename R, bool FIXED>
__global__ void __launch_bounds__(PME_THREADS) k_pmeSpread(PmeArgs a) {
    __shared__ PmeAtom<R> sAtom[PME_ATOMS];
    if (threadIdx.x < PME_ATOMS) {
        const int k = blockIdx.x*PME_ATOMS+threadIdx.x;
        PmeAtom<R>& me = sAtom[threadIdx.x];
        me.subset = -1;
        if (k < a.numAtoms) {
            const int atom = a.sortedAtom ? a.sortedAtom[k] : k;
            if (atom >= 0) {
                const R q = pmeCharge<R>(a, atom);
                if (q != 0) {
                    R frac[3], dt[ORDER];
                    gridCoord<R>(a, atom, me.idx, frac);
                    bspline5<R>(frac[0], me.th[0], dt);
                    bspline5<R>(frac[1], me.th[1], dt);
                    bspline5<R>(frac[2], me.th[2], dt);
                    me.q = q;
                    me.subset = __float_as_int(a.params4[atom].w);
                }
            }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x&31, warp = threadIdx.x>>5;
    if (lane >= ORDER*ORDER)
        return;
    const int iy = lane/ORDER, iz = lane-iy*ORDER;
    for (int i = 0; i < PME_ATOMS_PER_WARP; i++) {
        const PmeAtom<R>& at = sAtom[warp*PME_ATOMS_PER_WARP+i];
        if (at.subset < 0)
            continue;
        int y = at.idx[1]+iy, z = at.idx[2]+iz;
        y -= y >= a.ny ? a.ny : 0;
        z -= z >= a.nz ? a.nz : 0;
        const R wyz = at.q*at.th[1][iy]*at.th[2][iz];
        const size_t plane = (size_t) a.ny*a.nz;
        const size_t base = (size_t) at.subset*a.nx*plane+(size_t) y*a.nz+z;
#pragma unroll
        for (int ix = 0; ix < ORDER; ix++) {
            int x = at.idx[0]+ix;
            x -= x >= a.nx ? a.nx : 0;
            const R v = wyz*at.th[0][ix];
            if (FIXED)
                atomicAdd((unsigned long long*) a.gridFixed+base+x*plane, (unsigned long long) __double2ll_rn((double) v*PME_FIXED_SCALE));
            else
                atomicAdd((R*) a.grid+base+x*plane, v);
        }
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

// ---- potentials seen by each subset: phi_i = sum_j lambda[slice(i,j)] * grid_j, in place, so that the
// interpolation reads ONE value per stencil point whatever the number of subsets (the reference applies
// the lambdas inside the interpolation loop, ReferencePME.cpp:634-701; same sums, reordered).
template <typename R>
__global__ void __launch_bounds__(256) k_pmeCombine(PmeArgs a, size_t G) {
    __shared__ R sLambda[SNB_MAX_SUBSETS*SNB_MAX_SUBSETS];
    const int S = a.numSubsets;
    if (threadIdx.x < S*S) {
        const int si = threadIdx.x/S, sj = threadIdx.x-si*S;
        sLambda[threadIdx.x] = (R) (a.term == 0 ? a.sp->lambdaC : a.sp->lambdaV)[sliceOf(si, sj)];
    }
    __syncthreads();
    R* grid = (R*) a.grid;
    for (size_t i = blockIdx.x*(size_t) blockDim.x+threadIdx.x; i < G; i += (size_t) gridDim.x*blockDim.x) {
        R v[SNB_MAX_SUBSETS];
        for (int j = 0; j < S; j++)
            v[j] = grid[j*G+i];
        for (int k = 0; k < S; k++) {
            R phi = 0;
            for (int j = 0; j < S; j++)
                phi += sLambda[k*S+j]*v[j];
            grid[k*G+i] = phi;
        }
    }
}

// ---- interpolation: same two phases; phase 2 gathers 5 x planes of the 5x5 yz face per lane, for every
// subset's potential grid weighted by the lambda of the slice it forms with the atom's subset, and reduces
// the three force components over the 25 lanes by warp shuffle.
template <typename R> struct PmeAtomD {
    R th[3][ORDER], dth[3][ORDER];
    R q;
    int idx[3];
    int subset, atom;   // subset -1: nothing to do
    R f[3];             // phase 2 result: sums over the stencil of (dtheta theta theta, ...) * phi
    R lam[2];           // FOLD: lambda of the slices the atom's subset forms with subsets 0 and 1
};

// FOLD = 0: the grids were lambda-combined by k_pmeCombine; FOLD = 1 / 2 (one / two subsets): the combination is done here,
// on the 125 points an atom reads (the same multiply-add sequence as k_pmeCombine: bit-identical potentials) -- the
// combine pass over the whole grid and its launch drop out of the reciprocal-space chain.
template <typename R, int FOLD>
__global__ void __launch_bounds__(PME_THREADS) k_pmeInterpolate(PmeArgs a) {
    __shared__ PmeAtomD<R> sAtom[PME_ATOMS];
    if (threadIdx.x < PME_ATOMS) {
        const int k = blockIdx.x*PME_ATOMS+threadIdx.x;
        PmeAtomD<R>& me = sAtom[threadIdx.x];
        me.subset = -1;
        me.idx[0] = me.idx[1] = me.idx[2] = 0;      // (phase 2 gathers without a branch: a valid stencil for every slot)
        if (k < a.numAtoms) {
            const int atom = a.sortedAtom ? a.sortedAtom[k] : k;
            if (atom >= 0) {
                const R q = pmeCharge<R>(a, atom);
                if (q != 0) {
                    R frac[3];
                    gridCoord<R>(a, atom, me.idx, frac);
                    bspline5<R>(frac[0], me.th[0], me.dth[0]);
                    bspline5<R>(frac[1], me.th[1], me.dth[1]);
                    bspline5<R>(frac[2], me.th[2], me.dth[2]);
                    me.q = q;
                    me.subset = __float_as_int(a.params4[atom].w);
                    me.atom = atom;
                    if (FOLD) {
                        const double* lambda = a.term == 0 ? a.sp->lambdaC : a.sp->lambdaV;
                        me.lam[0] = (R) lambda[sliceOf(me.subset, 0)];
                        me.lam[1] = FOLD > 1 ? (R) lambda[sliceOf(me.subset, 1)] : (R) 0;
                    }
                }
            }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x&31, warp = threadIdx.x>>5;
    const bool active = lane < ORDER*ORDER;
    const int iy = active ? lane/ORDER : 0, iz = active ? lane-iy*ORDER : 0;
    const size_t plane = (size_t) a.ny*a.nz, G = plane*a.nx;
    // (the warp's atoms four at a time: their 20 gathers per lane are in flight together and the four reductions interleave --
    // one atom at a time was eight L2 round trips and eight chains of shuffles in a row)
    constexpr int BATCH = 4;
    static_assert(PME_ATOMS_PER_WARP%BATCH == 0, "interpolation batches");
    for (int i0 = 0; i0 < PME_ATOMS_PER_WARP; i0 += BATCH) {
        R fx[BATCH], fy[BATCH], fz[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const PmeAtomD<R>& at = sAtom[warp*PME_ATOMS_PER_WARP+i0+u];
            // no branch around the gathers (a guarded block per atom would serialise them again): an idle lane or an atom with
            // nothing to do reads the stencil of grid point (0, 0, 0) of subset 0 and its result is dropped
            const bool on = active && at.subset >= 0;
            const int sub = max(at.subset, 0);
            int y = at.idx[1]+iy, z = at.idx[2]+iz;
            y -= y >= a.ny ? a.ny : 0;
            z -= z >= a.nz ? a.nz : 0;
            const R* __restrict__ cell = (const R*) a.grid+(FOLD ? 0 : sub*G)+(size_t) y*a.nz+z;
            R phi[ORDER];
            if (FOLD) {
                R v0[ORDER], v1[ORDER];
#pragma unroll
                for (int ix = 0; ix < ORDER; ix++) {
                    int x = at.idx[0]+ix;
                    x -= x >= a.nx ? a.nx : 0;
                    v0[ix] = __ldg(cell+x*plane);
                    v1[ix] = FOLD > 1 ? __ldg(cell+G+x*plane) : (R) 0;
                }
#pragma unroll
                for (int ix = 0; ix < ORDER; ix++) {
                    phi[ix] = at.lam[0]*v0[ix];
                    if (FOLD > 1)
                        phi[ix] += at.lam[1]*v1[ix];
                }
            }
            else {
#pragma unroll
                for (int ix = 0; ix < ORDER; ix++) {
                    int x = at.idx[0]+ix;
                    x -= x >= a.nx ? a.nx : 0;
                    phi[ix] = __ldg(cell+x*plane);      // already lambda-combined for this subset (k_pmeCombine)
                }
            }
            R s0 = 0, s1 = 0;               // sum_x theta_x * phi, sum_x dtheta_x * phi
#pragma unroll
            for (int ix = 0; ix < ORDER; ix++) {
                s0 += at.th[0][ix]*phi[ix];
                s1 += at.dth[0][ix]*phi[ix];
            }
            fx[u] = on ? s1*at.th[1][iy]*at.th[2][iz] : (R) 0;
            fy[u] = on ? s0*at.dth[1][iy]*at.th[2][iz] : (R) 0;
            fz[u] = on ? s0*at.th[1][iy]*at.dth[2][iz] : (R) 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < BATCH; u++) {
                fx[u] += __shfl_xor_sync(0xffffffffu, fx[u], o);
                fy[u] += __shfl_xor_sync(0xffffffffu, fy[u], o);
                fz[u] += __shfl_xor_sync(0xffffffffu, fz[u], o);
            }
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < BATCH; u++) {
                PmeAtomD<R>& at = sAtom[warp*PME_ATOMS_PER_WARP+i0+u];
                at.f[0] = fx[u]; at.f[1] = fy[u]; at.f[2] = fz[u];
            }
        }
    }
    __syncthreads();
    // phase 3, one thread per atom: to Cartesian forces (ReferencePME.cpp:698-700), in double
    if (threadIdx.x < PME_ATOMS) {
        const PmeAtomD<R>& at = sAtom[threadIdx.x];
        if (at.subset >= 0) {
            const BoxD boxd = a.sp->exact.boxd;
            const double qd = at.q;
            const double gx = (double) at.f[0]*a.nx, gy = (double) at.f[1]*a.ny, gz = (double) at.f[2]*a.nz;
            const double Fx = -qd*(gx*boxd.recip[0][0]);
            const double Fy = -qd*(gx*boxd.recip[1][0]+gy*boxd.recip[1][1]);
            const double Fz = -qd*(gx*boxd.recip[2][0]+gy*boxd.recip[2][1]+gz*boxd.recip[2][2]);
            addFixed(&a.forceOrig[at.atom], toFixedD(Fx));
            addFixed(&a.forceOrig[a.paddedAtoms+at.atom], toFixedD(Fy));
            addFixed(&a.forceOrig[2*a.paddedAtoms+at.atom], toFixedD(Fz));
        }
    }
}

void launchPmeEterm(const PmeArgs& a, cudaStream_t stream) {
    if (a.doublePrecision) k_pmeEterm<double><<<592, 256, 0, stream>>>(a, (double*) a.eterm);
    else k_pmeEterm<float><<<592, 256, 0, stream>>>(a, (float*) a.eterm);
}
// the grid the spreading accumulates into, zeroed: part of launchPmeSpread, or issued ahead of it (beside the sort, which
// the spreading waits for) with zeroed = true there
void launchPmeZero(const PmeArgs& a, cudaStream_t stream) {
    const size_t total = (size_t) a.numSubsets*a.nx*a.ny*a.nz;
    if (a.gridFixed) cudaMemsetAsync(a.gridFixed, 0, sizeof(long long)*total, stream);
    else cudaMemsetAsync(a.grid, 0, (a.doublePrecision ? 8 : 4)*total, stream);
}
void launchPmeSpread(const PmeArgs& a, cudaStream_t stream, bool zeroed) {
    const size_t total = (size_t) a.numSubsets*a.nx*a.ny*a.nz;
    const int blocks = (a.numAtoms+PME_ATOMS-1)/PME_ATOMS;
    if (!zeroed)
        launchPmeZero(a, stream);
    if (a.gridFixed) {
        int fb = (int) std::min<size_t>((total+255)/256, 2368);
        if (a.doublePrecision) {
            k_pmeSpread<double, true><<<blocks, PME_THREADS, 0, stream>>>(a);
            k_pmeFinishSpread<double><<<fb, 256, 0, stream>>>(a, total);
        }
        else {
            k_pmeSpread<float, true><<<blocks, PME_THREADS, 0, stream>>>(a);
            k_pmeFinishSpread<float><<<fb, 256, 0, stream>>>(a, total);
        }
    }
    else {
        if (a.doublePrecision) k_pmeSpread<double, false><<<blocks, PME_THREADS, 0, stream>>>(a);
        else k_pmeSpread<float, false><<<blocks, PME_THREADS, 0, stream>>>(a);
    }
}
void launchPmeConvolution(const PmeArgs& a, cudaStream_t stream) {
    size_t total = (size_t) a.nx*a.ny*(a.nz/2+1);
    int blocks = (int) std::min<size_t>((total+255)/256, 1184);
    if (a.doublePrecision) k_pmeConvolution<double><<<blocks, 256, 0, stream>>>(a);
    else k_pmeConvolution<float><<<blocks, 256, 0, stream>>>(a);
}
bool pmeInterpolateFolds(const PmeArgs& a) {
    const char* env = getenv("SNB_PME_FOLD");       // (read at enqueue / capture time only)
    return a.numSubsets <= 2 && !(env && atoi(env) == 0);
}
void launchPmeInterpolate(const PmeArgs& a, cudaStream_t stream) {
    const int blocks = (a.numAtoms+PME_ATOMS-1)/PME_ATOMS;
    const size_t G = (size_t) a.nx*a.ny*a.nz;
    const int cb = (int) std::min<size_t>((G+255)/256, 2368);
    // one or two subsets: the lambda combination happens inside the interpolation (SNB_PME_FOLD=0: a pass of its own, as for
    // more subsets)
    if (pmeInterpolateFolds(a)) {
        if (a.numSubsets == 1) {
            if (a.doublePrecision) k_pmeInterpolate<double, 1><<<blocks, PME_THREADS, 0, stream>>>(a);
            else k_pmeInterpolate<float, 1><<<blocks, PME_THREADS, 0, stream>>>(a);
        }
        else {
            if (a.doublePrecision) k_pmeInterpolate<double, 2><<<blocks, PME_THREADS, 0, stream>>>(a);
            else k_pmeInterpolate<float, 2><<<blocks, PME_THREADS, 0, stream>>>(a);
        }
        return;
    }
    if (a.doublePrecision) k_pmeCombine<double><<<cb, 256, 0, stream>>>(a, G);
    else k_pmeCombine<float><<<cb, 256, 0, stream>>>(a, G);
    if (a.doublePrecision) k_pmeInterpolate<double, 0><<<blocks, PME_THREADS, 0, stream>>>(a);
    else k_pmeInterpolate<float, 0><<<blocks, PME_THREADS, 0, stream>>>(a);
}
