// thb_stft_fast.cu -- register-resident warp-per-frame path (placeholder until the fast kernel lands).
#include "thb_kernels.cuh"

namespace thb {

bool stft_fast_supported(const PlanDev &) { return false; }

cudaError_t launch_stft_fast(const PlanDev &, const TrackDesc *, int, long long, int, cudaStream_t) {
    return cudaErrorNotSupported;
}

}  // namespace thb
