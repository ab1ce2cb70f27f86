// Final pass: gather the two fixed-point force accumulators (sorted-order tile forces and
// original-order bonded / reciprocal forces) into the caller's layout, and the small helper that
// widens device float4 positions to the double array the exact predicates read.
#include "kernels.cuh"

__global__ void k_finish(FinishArgs a) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (a.hostEnergy) {
        // every producer of energies (direct space, exceptions, reciprocal space) was joined before this kernel
        for (int k = i; k < a.numEnergyHost; k += gridDim.x*blockDim.x)
            a.hostEnergy[k] = a.energySrc[k];
        if (i < 8)
            a.hostCounters[i] = a.counters[i];
    }
    if (a.reduceOut && i < a.numEnergy) {
        const double v = i == a.overflowSlot ? (double) a.counters[2] : a.energyIn[i];
        a.reduceOut[3*(size_t) a.numAtoms+i] = __double2ll_rn(v*SNB_FORCE_SCALE_D);
    }
    if (i >= a.numAtoms)
        return;
    // thread t serves the atom of the caller's slot t (original order unless the caller handed its own)
    const int t = i;
    if (a.outIndex)
        i = a.outIndex[t];
    const int k = a.sortedPos ? a.sortedPos[i] : i;
    // a pass whose tile list overflowed is discarded and repeated by the host: it must leave the caller's accumulator alone
    const bool discard = a.counters[2] != 0;
    for (int c = 0; c < 3; c++) {
        long long f = a.forceSorted[c*a.paddedAtoms+k]+a.forceOrig[c*a.paddedAtoms+i];
        if (a.forcesOut)
            a.forcesOut[3*i+c] = (double) f*(1.0/SNB_FORCE_SCALE_D);
        if (a.forcesFixedOut && !discard)
            a.forcesFixedOut[(size_t) c*a.outStride+t] += f;
        if (a.reduceOut)
            a.reduceOut[(size_t) c*a.numAtoms+i] = f;
    }
}

__global__ void k_emitReduced(const long long* reduced, int numAtoms, int paddedAtoms, const long long* extra, double* forcesOut, long long* forcesFixedOut,
                              const int* outIndex, int outStride, const long long* overflow) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (i >= numAtoms)
        return;
    const int t = i;
    if (outIndex)
        i = outIndex[t];
    // some rank's tile list overflowed: every rank repeats the evaluation, nothing of this pass may reach the caller's accumulator
    const bool discard = overflow && *overflow != 0;
    for (int c = 0; c < 3; c++) {
        // extra: rank 0's reciprocal-space forces ([3][paddedAtoms], original order), computed while the reduction ran
        const long long f = reduced[(size_t) c*numAtoms+i]+(extra ? extra[(size_t) c*paddedAtoms+i] : 0ll);
        if (forcesOut)
            forcesOut[3*i+c] = (double) f*(1.0/SNB_FORCE_SCALE_D);
        if (forcesFixedOut && !discard)
            forcesFixedOut[(size_t) c*outStride+t] += f;
    }
}

__global__ void k_convertPositions(const void* posq, const float4* correction, const int* atomIndex, int positionsDouble, double* posd, int numAtoms) {
    int t = blockIdx.x*blockDim.x+threadIdx.x;
    if (t >= numAtoms)
        return;
    double x, y, z;
    if (positionsDouble) {
        const double4 p = ((const double4*) posq)[t];
        x = p.x; y = p.y; z = p.z;
    }
    else {
        const float4 p = ((const float4*) posq)[t];
        x = p.x; y = p.y; z = p.z;
        if (correction) {
            const float4 c = correction[t];
            x += (double) c.x; y += (double) c.y; z += (double) c.z;
        }
    }
    const int i = atomIndex ? atomIndex[t] : t;
    posd[3*i] = x; posd[3*i+1] = y; posd[3*i+2] = z;
}

__global__ void __launch_bounds__(256) k_begin(BeginArgs a) {
    const size_t tid = blockIdx.x*(size_t) blockDim.x+threadIdx.x, nThreads = gridDim.x*(size_t) blockDim.x;
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < a.stepWords; k += blockDim.x)
            a.stepDev[k] = ((const volatile int*) a.stepHost)[k];
    for (size_t k = tid; k < a.zeroWords; k += nThreads)
        a.zero[k] = 0ull;
    if (a.posq)
        for (size_t t = tid; t < (size_t) a.numAtoms; t += nThreads) {
            double x, y, z;
            if (a.positionsDouble) {
                const double4 p = ((const double4*) a.posq)[t];
                x = p.x; y = p.y; z = p.z;
            }
            else {
                const float4 p = ((const float4*) a.posq)[t];
                x = p.x; y = p.y; z = p.z;
                if (a.correction) {
                    const float4 c = a.correction[t];
                    x += (double) c.x; y += (double) c.y; z += (double) c.z;
                }
            }
            const size_t i = a.atomIndex ? (size_t) a.atomIndex[t] : t;
            a.posd[3*i] = x; a.posd[3*i+1] = y; a.posd[3*i+2] = z;
        }
}

void launchBegin(const BeginArgs& a, cudaStream_t stream) {
    const size_t work = a.zeroWords/4 > (size_t) a.numAtoms ? a.zeroWords/4 : (size_t) a.numAtoms;
    const unsigned grid = (unsigned) ((work+255)/256 > 0 ? (work+255)/256 : 1);
    k_begin<<<grid, 256, 0, stream>>>(a);
}

void launchFinish(const FinishArgs& a, cudaStream_t stream) {
    const int n = a.reduceOut ? (a.numAtoms > a.numEnergy ? a.numAtoms : a.numEnergy) : a.numAtoms;
    if (n > 0)
        k_finish<<<(n+255)/256, 256, 0, stream>>>(a);
}

void launchEmitReduced(const long long* reduced, int numAtoms, int paddedAtoms, const long long* extra, double* forcesOut, long long* forcesFixedOut,
                       const int* outIndex, int outStride, const long long* overflow, cudaStream_t stream) {
    if (numAtoms > 0)
        k_emitReduced<<<(numAtoms+255)/256, 256, 0, stream>>>(reduced, numAtoms, paddedAtoms, extra, forcesOut, forcesFixedOut, outIndex, outStride, overflow);
}

void launchConvertPositions(const void* posq, const void* correction, const int* atomIndex, int positionsDouble, double* posd, int numAtoms, cudaStream_t stream) {
    if (numAtoms > 0)
        k_convertPositions<<<(numAtoms+255)/256, 256, 0, stream>>>(posq, (const float4*) correction, atomIndex, positionsDouble, posd, numAtoms);
}
