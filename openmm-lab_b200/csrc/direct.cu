// Dispatch of the direct-space tile kernel over interaction kinds (kernels live in direct_impl.cuh).
#include "kernels.cuh"

#include <cstdlib>

template <int KIND> void launchDirectKind(const DirectArgs& a, cudaStream_t st, int grid);
bool launchDirectPacked(const DirectArgs& a, cudaStream_t st, int grid);    // direct_packed.cu

// returns true when the packed-FP32 kernel took the launch (reported through snb_get_counters)
bool launchDirect(const DirectArgs& a, cudaStream_t st, int numSMs) {
    int grid = numSMs*DIRECT_CTAS_PER_SM;       // one-warp CTAs, all resident, equal contiguous shares of the chunk list
    // A/B knob: SNB_DIRECT_WAVES=w launches w times as many CTAs (each with 1/w of the share), so that CTAs retire while the
    // kernel runs and the (higher-priority) kernels of reciprocal space find room beside it; a fraction leaves SMs part empty
    static const double waves = getenv("SNB_DIRECT_WAVES") ? atof(getenv("SNB_DIRECT_WAVES")) : 1.0;
    if (waves > 0 && waves != 1.0)
        grid = max(1, (int) (numSMs*waves+0.5))*DIRECT_CTAS_PER_SM;
    // PME direct space in an orthorhombic box with <= 2 subsets has two kernels: the packed-FP32 pair loop (large
    // systems) and the general one (api.cu decides: SNB_FLAG_DIRECT_*, else by size)
    if (a.usePacked && launchDirectPacked(a, st, grid))
        return true;
    switch (a.kind) {
        case KIND_PLAIN: launchDirectKind<KIND_PLAIN>(a, st, grid); break;
        case KIND_RF: launchDirectKind<KIND_RF>(a, st, grid); break;
        case KIND_EWALD: launchDirectKind<KIND_EWALD>(a, st, grid); break;
        default: launchDirectKind<KIND_LJPME>(a, st, grid); break;
    }
    return false;
}
