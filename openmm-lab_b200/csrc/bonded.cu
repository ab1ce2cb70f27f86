// Exception / exclusion-correction kernel (north-star subsystem 3).  One thread per exception.
//
//  * every exception pair (also the zeroed ones) gets the reciprocal-space image subtracted when
//    an Ewald-family method is in use: ReferenceSlicedLJCoulombIxn.cpp:433-490
//    (GPU behavioural cross-reference: platforms/common/src/kernels/pmeExclusions.cc:1-46);
//  * exceptions with chargeProd != 0, eps != 0 or an offset are 1-4 interactions:
//    ReferenceSlicedLJCoulomb14.cpp:61-95 (cross-reference nonbondedExceptions.cc:1-24).
//
// There are O(N) of these pairs and they sit at bonded distances where erf-vs-1/r cancellation is
// worst, so the whole kernel is FP64 (B200 FP64 is half the FP32 rate; the cost is noise next to
// the tile kernel) and reads the untouched double-precision input coordinates.
#include "kernels.cuh"

__global__ void __launch_bounds__(128) k_bonded(BondedArgs a) {
    __shared__ double sE[2*SNB_MAX_SLICES];
    for (int k = threadIdx.x; k < 2*SNB_MAX_SLICES; k += blockDim.x)
        sE[k] = 0.0;
    __syncthreads();
    const int idx = blockIdx.x*blockDim.x+threadIdx.x;
    int mySlice = -1;
    double myC = 0.0, myV = 0.0;
    if (idx < a.numExceptions) {
        const int e = a.firstException+idx;
        const int2 atoms = a.exceptionAtoms[e];
        const int i = atoms.x, j = atoms.y;
        const int slice = sliceOf(a.subset[i], a.subset[j]);
        // delta = pos[i] - pos[j]  (getDeltaR(pos[j], pos[i]))
        double d[3];
        deltaExact(a.posd+3*(size_t) j, a.posd+3*(size_t) i, a.sp->exact.boxd, a.periodicExceptions != 0, d);
        const double r2 = d[0]*d[0]+d[1]*d[1]+d[2]*d[2];
        const double r = sqrt(r2), invR = 1.0/r;
        double fi = 0.0;            // force on i is  +fi * d, on j  -fi * d
        double eC = 0.0, eV = 0.0;
        const double K = SNB_ONE_4PI_EPS0;
        if (a.ewald) {
            const double qq = K*a.paramsD[3*i+2]*a.paramsD[3*j+2];
            const double alphaR = a.alpha*r;
            const double erfAR = erf(alphaR);
            if (erfAR > 1e-6) {
                const double dEdR = qq*invR*invR*invR*(erfAR-2*alphaR*exp(-alphaR*alphaR)/sqrt(M_PI));
                fi -= a.sp->lambdaC[slice]*dEdR;
                eC -= qq*invR*erfAR;
            }
            else
                eC -= a.alpha*(2/sqrt(M_PI))*qq;
            if (a.ljpme) {
                const double dar2 = (a.dalpha*r)*(a.dalpha*r), dar4 = dar2*dar2, dar6 = dar4*dar2;
                const double invR2 = invR*invR, invR6 = invR2*invR2*invR2;
                const double si = a.paramsD[3*i], sj = a.paramsD[3*j];
                const double c6i = 8.0*si*si*si*a.paramsD[3*i+1], c6j = 8.0*sj*sj*sj*a.paramsD[3*j+1];
                const double ex = exp(-dar2);
                eV += c6i*c6j*invR6*(1.0-ex*(1.0+dar2+0.5*dar4));
                const double dEdR = -6.0*c6i*c6j*invR6*invR2*(1.0-ex*(1.0+dar2+0.5*dar4+dar6/6.0));
                fi -= a.sp->lambdaV[slice]*dEdR;
            }
        }
        if (a.exceptionIs14[e]) {
            const double sig = a.exception14[3*e], eps4 = a.exception14[3*e+1], qq = a.exception14[3*e+2];
            double sig2 = invR*sig;
            sig2 *= sig2;
            const double sig6 = sig2*sig2*sig2;
            double dEdR = a.sp->lambdaV[slice]*eps4*(12.0*sig6-6.0)*sig6;
            dEdR += a.sp->lambdaC[slice]*K*qq*invR;
            dEdR *= invR*invR;
            fi += dEdR;
            eV += eps4*(sig6-1.0)*sig6;
            eC += K*qq*invR;
        }
        if (a.includeForces && fi != 0.0) {
            for (int k = 0; k < 3; k++) {
                long long f = toFixedD(fi*d[k]);
                addFixed(&a.forceOrig[k*a.paddedAtoms+i], f);
                addFixed(&a.forceOrig[k*a.paddedAtoms+j], -f);
            }
        }
        mySlice = slice;
        myC = eC; myV = eV;
    }
    // Energies: summed over the warp slice by slice (a warp's exceptions sit in one or two slices) before they touch shared
    // memory -- a double-precision atomicAdd there is a compare-and-swap loop, and with every thread of the CTA adding to the
    // same two words it WAS the kernel (168 us for the two million exceptions of the STMV-size system) -- then one global
    // atomic per CTA and slice into one of the 32 replicas of the table (summed on the host)
    {
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x&31;
        unsigned todo = __ballot_sync(full, mySlice >= 0);
        while (todo) {
            const int s = __shfl_sync(full, mySlice, __ffs(todo)-1);
            const bool mine = mySlice == s;
            double c = mine ? myC : 0.0, v = mine ? myV : 0.0;
            for (int o = 16; o > 0; o >>= 1) {
                c += __shfl_xor_sync(full, c, o);
                v += __shfl_xor_sync(full, v, o);
            }
            if (lane == 0) {
                if (c != 0.0) atomicAdd(&sE[2*s], c);
                if (v != 0.0) atomicAdd(&sE[2*s+1], v);
            }
            todo &= ~__ballot_sync(full, mine);
        }
    }
    __syncthreads();
    double* table = a.energyBuf+(size_t) (blockIdx.x%32)*2*SNB_MAX_SLICES;
    for (int k = threadIdx.x; k < 2*SNB_MAX_SLICES; k += blockDim.x)
        if (sE[k] != 0.0)
            atomicAdd(&table[k], sE[k]);
}

void launchBonded(const BondedArgs& a, cudaStream_t stream) {
    if (a.numExceptions == 0)
        return;
    k_bonded<<<(a.numExceptions+127)/128, 128, 0, stream>>>(a);
}
