// thb_stft2048.cuh -- pieces shared by the two n_fft == 2048 kernels (thb_stft_fast.cu: one frame per warp,
// scalar; thb_stft_pair.cu: two frames per warp, packed f32x2).  Both follow the same operation order, so their
// results agree bit for bit and a frame may be computed by either.
#pragma once
#include "thb_device.cuh"
#include "thb_kernels.cuh"

namespace thb {
namespace k2048 {

constexpr int kRow = 33;      // row stride of the 32 x 32 transpose tile (conflict-free both ways)
constexpr int kMagBase = 16;  // the mel walk may start up to 15 bins before bin 0

// cos / sin of 2 pi j / 32
__device__ constexpr float kC32[32] = {
    1.0f, 0.9807852506637573f, 0.9238795042037964f, 0.8314695954322815f, 0.7071067690849304f, 0.5555702447891235f,
    0.3826834261417389f, 0.19509032368659973f, 0.0f, -0.19509032368659973f, -0.3826834261417389f, -0.5555702447891235f,
    -0.7071067690849304f, -0.8314695954322815f, -0.9238795042037964f, -0.9807852506637573f, -1.0f,
    -0.9807852506637573f, -0.9238795042037964f, -0.8314695954322815f, -0.7071067690849304f, -0.5555702447891235f,
    -0.3826834261417389f, -0.19509032368659973f, 0.0f, 0.19509032368659973f, 0.3826834261417389f, 0.5555702447891235f,
    0.7071067690849304f, 0.8314695954322815f, 0.9238795042037964f, 0.9807852506637573f};
__device__ constexpr float kS32[32] = {
    0.0f, 0.19509032368659973f, 0.3826834261417389f, 0.5555702447891235f, 0.7071067690849304f, 0.8314695954322815f,
    0.9238795042037964f, 0.9807852506637573f, 1.0f, 0.9807852506637573f, 0.9238795042037964f, 0.8314695954322815f,
    0.7071067690849304f, 0.5555702447891235f, 0.3826834261417389f, 0.19509032368659973f, 0.0f, -0.19509032368659973f,
    -0.3826834261417389f, -0.5555702447891235f, -0.7071067690849304f, -0.8314695954322815f, -0.9238795042037964f,
    -0.9807852506637573f, -1.0f, -0.9807852506637573f, -0.9238795042037964f, -0.8314695954322815f,
    -0.7071067690849304f, -0.5555702447891235f, -0.3826834261417389f, -0.19509032368659973f};

// position of output X[k] of the in-register 32-point DFT (radix 4 x 8, outputs left where they fall)
__device__ __forceinline__ constexpr int perm32(int k) { return 8 * (k & 3) + (k >> 2); }

__device__ __forceinline__ float lg2_ftz(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_ftz(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr float kDbPerLog2Pow = 3.01029995663981195f;   // 10 log10(2): dB per doubling of power
constexpr float kDbPerLog2Amp = 6.02059991327962390f;   // 20 log10(2)
// A frame whose largest |X|^2 leaves [2^-80, 2^100] (f32 audio hundreds of dB from full scale) would lose bins
// to under/overflow of the square.  The scalar kernel re-runs such a frame once on a spectrum rescaled by
// 2^(+-60) (exact) and shifts the dB back; the frame-pair kernel hands the tile to the scalar kernel.
constexpr float kPowTiny = 8.2718061e-25f;     // 2^-80
constexpr float kPowHuge = 1.2676506e+30f;     // 2^100
constexpr float kRescueUp = 1.1529215e+18f;    // 2^60
constexpr float kRescueDown = 8.6736174e-19f;  // 2^-60

// view of MelItems::blob() (thb_host.hpp) once it sits in shared memory
struct MelView {
    const uint32_t *base;
    int n_groups, n_mel;
    const uint32_t *T, *woff, *pptr, *pids;
    const int32_t *start;
    __device__ __forceinline__ explicit MelView(const uint32_t *b) : base(b) {
        n_groups = static_cast<int>(b[0]);
        n_mel = static_cast<int>(b[1]);
        T = b + b[2];
        woff = b + b[3];
        start = reinterpret_cast<const int32_t *>(b + b[4]);
        pptr = b + b[5];
        pids = b + b[6];
    }
};

// elements (float for the scalar kernel, float2 for the pair kernel) of one warp's tile: the transpose tile,
// later the magnitudes [kMagBase + bin] with the mel walk's lead / reach, then one partial sum per mel slot
__host__ __device__ inline int part_base(const PlanDev &p) { return (kMagBase + (p.n_mel ? p.mi_max_reach : 1024) + 2) & ~1; }
__host__ __device__ inline int tile_elems(const PlanDev &p) {
    const int need = part_base(p) + (p.n_mel ? p.mi_groups * 32 : 0);
    const int t = need > 32 * kRow ? need : 32 * kRow;
    return (t + 3) & ~3;
}

}  // namespace k2048
}  // namespace thb
