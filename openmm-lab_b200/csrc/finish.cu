// Final pass: gather the two fixed-point force accumulators (sorted-order tile forces and
// original-order bonded / reciprocal forces) into the caller's layout, and the small helper that
// widens device float4 positions to the double array the exact predicates read.
#include "kernels.cuh"

__global__ void k_finish(FinishArgs a) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (a.reduceOut && i < a.numEnergy) {
        const double v = i == a.overflowSlot ? (double) a.counters[2] : a.energyIn[i];
        a.reduceOut[3*(size_t) a.numAtoms+i] = __double2ll_rn(v*SNB_FORCE_SCALE_D);
    }
    if (i >= a.numAtoms)
        return;
    const int k = a.sortedPos ? a.sortedPos[i] : i;
    for (int c = 0; c < 3; c++) {
        long long f = a.forceSorted[c*a.paddedAtoms+k]+a.forceOrig[c*a.paddedAtoms+i];
        if (a.forcesOut)
            a.forcesOut[3*i+c] = (double) f*(1.0/SNB_FORCE_SCALE_D);
        if (a.forcesFixedOut)
            a.forcesFixedOut[(size_t) c*a.numAtoms+i] += f;
        if (a.reduceOut)
            a.reduceOut[(size_t) c*a.numAtoms+i] = f;
    }
}

__global__ void k_emitReduced(const long long* reduced, int numAtoms, double* forcesOut, long long* forcesFixedOut) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (i >= numAtoms)
        return;
    for (int c = 0; c < 3; c++) {
        const long long f = reduced[(size_t) c*numAtoms+i];
        if (forcesOut)
            forcesOut[3*i+c] = (double) f*(1.0/SNB_FORCE_SCALE_D);
        if (forcesFixedOut)
            forcesFixedOut[(size_t) c*numAtoms+i] += f;
    }
}

__global__ void k_convertPositions(const float4* posq, double* posd, int numAtoms) {
    int i = blockIdx.x*blockDim.x+threadIdx.x;
    if (i >= numAtoms)
        return;
    float4 p = posq[i];
    posd[3*i] = p.x; posd[3*i+1] = p.y; posd[3*i+2] = p.z;
}

void launchFinish(const FinishArgs& a, cudaStream_t stream) {
    const int n = a.reduceOut ? (a.numAtoms > a.numEnergy ? a.numAtoms : a.numEnergy) : a.numAtoms;
    if (n > 0)
        k_finish<<<(n+255)/256, 256, 0, stream>>>(a);
}

void launchEmitReduced(const long long* reduced, int numAtoms, double* forcesOut, long long* forcesFixedOut, cudaStream_t stream) {
    if (numAtoms > 0)
        k_emitReduced<<<(numAtoms+255)/256, 256, 0, stream>>>(reduced, numAtoms, forcesOut, forcesFixedOut);
}

void launchConvertPositions(const float4* posq, double* posd, int numAtoms, cudaStream_t stream) {
    if (numAtoms > 0)
        k_convertPositions<<<(numAtoms+255)/256, 256, 0, stream>>>(posq, posd, numAtoms);
}
