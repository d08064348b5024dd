// thb_stft_warp.cu -- K1/K2/K3 for n_fft == 1024 and n_fft == 512: the warp-register design of the n_fft == 2048
// kernels (thb_stft_pair.cu / thb_stft_fast.cu) for the reference's default setting at 8 - 24 kHz
// (SpecSetting::default = 40 ms / 4: win 320 .. 960, n_fft 512 / 1024; mod.rs:238-274 are its own fixture rates).
//
//   frame  : n_fft = 64 R1 real samples packed as NC = 32 R1 complex z[m] = x[2m] + i x[2m+1], R1 = 16 | 8.
//   warp   : G = 32 / R1 frame GROUPS at once, R1 lanes each.  Pass 1: lane n2 holds z[32 n1 + n2], n1 < R1, of
//            every group (coalesced 8-byte loads x window), one R1-point DFT per group in registers, twiddle
//            W_NC^(n2 k1), one 32 x 33 transpose through the warp's tile -- row g R1 + k1.  Pass 2: lane (g, k1)
//            owns row k1 of group g and runs the 32-point DFT: Z[k1 + R1 k2].  So the registers (32 complex
//            values per lane), the tile and both passes are exactly as full as in the n_fft == 2048 kernels.
//   split  : X[k], X[NC - k] from Z[k], Z[NC - k]; the partner sits in lane (R1 - k1) % R1 of the same group.
//   |X|, mel (bin-major MelItems schedule), dB through MUFU.LG2, running {max, -min}: as in the 2048 kernels.
//
// ONE kernel template, two value types: V = float2 holds the same quantity of two consecutive frames in every
// register pair and issues FADD2 / FMUL2 / FFMA2 (a "group" is a frame pair: 2 G frames per warp step); V = float is
// its scalar twin (one frame per group).  Both instantiate the same expressions in the same order, so their results
// agree bit for bit and a frame may be computed by either: the packed kernel takes the interior frames, the scalar
// one the file edges (reflect padding, utils.rs:111-137), unaligned leftovers and the tiles on the rescue list
// (frames whose |X|^2 leaves the exact range of the f32 square are redone on an exactly rescaled spectrum).
//
//   perform_stft (stft.rs:16-149), Complex::norm (spectrogram.rs:200), linspec.dot(mel_fb) (spectrogram.rs:207),
//   dB_from_amp (decibel.rs:198-202), find_min_max (mod.rs:169-178).
#include "thb_packed.cuh"

#include <cstdlib>

namespace thb {

namespace {

using namespace k2048;
using packed::kC16;
using packed::kS16;

// ---- arithmetic on V = float (one frame) or float2 (two frames, packed) ----
template <typename V>
struct Ops;
template <>
struct Ops<float> {
    static constexpr int kFrames = 1;
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float neg(float a) { return -a; }
    static __device__ __forceinline__ float muls(float a, float s) { return a * s; }
    static __device__ __forceinline__ float fmas(float a, float s, float c) { return fmaf(a, s, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float zero() { return 0.0f; }
};
template <>
struct Ops<float2> {
    static constexpr int kFrames = 2;
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ float2 sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
    static __device__ __forceinline__ float2 neg(float2 a) { return make_float2(-a.x, -a.y); }
    static __device__ __forceinline__ float2 muls(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
    static __device__ __forceinline__ float2 fmas(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
    static __device__ __forceinline__ float2 mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
    static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
    static __device__ __forceinline__ float2 zero() { return make_float2(0.0f, 0.0f); }
};

template <typename V>
struct Cx {
    V re, im;
};
template <typename V>
__device__ __forceinline__ Cx<V> cadd_t(Cx<V> a, Cx<V> b) { return {Ops<V>::add(a.re, b.re), Ops<V>::add(a.im, b.im)}; }
template <typename V>
__device__ __forceinline__ Cx<V> csub_t(Cx<V> a, Cx<V> b) { return {Ops<V>::sub(a.re, b.re), Ops<V>::sub(a.im, b.im)}; }
template <typename V>
__device__ __forceinline__ Cx<V> cmul_mi_t(Cx<V> a) { return {a.im, Ops<V>::neg(a.re)}; }  // * -i
// a * (c + i s)
template <typename V>
__device__ __forceinline__ Cx<V> cmul_cs(Cx<V> a, float c, float s) {
    Cx<V> r;
    r.re = Ops<V>::fmas(a.im, -s, Ops<V>::muls(a.re, c));
    r.im = Ops<V>::fmas(a.re, s, Ops<V>::muls(a.im, c));
    return r;
}

template <typename V>
__device__ __forceinline__ void dft4_t(Cx<V> &a0, Cx<V> &a1, Cx<V> &a2, Cx<V> &a3) {
    const Cx<V> s02 = cadd_t(a0, a2), d02 = csub_t(a0, a2);
    const Cx<V> s13 = cadd_t(a1, a3), d13 = cmul_mi_t(csub_t(a1, a3));
    a0 = cadd_t(s02, s13);
    a1 = cadd_t(d02, d13);
    a2 = csub_t(s02, s13);
    a3 = csub_t(d02, d13);
}

// natural-order 8-point DFT
template <typename V>
__device__ __forceinline__ void dft8_t(Cx<V> &v0, Cx<V> &v1, Cx<V> &v2, Cx<V> &v3, Cx<V> &v4, Cx<V> &v5, Cx<V> &v6, Cx<V> &v7) {
    using O = Ops<V>;
    const float h = 0.70710678118654752440f;
    Cx<V> t0 = cadd_t(v0, v4), u0 = csub_t(v0, v4);
    Cx<V> t1 = cadd_t(v1, v5), u1 = csub_t(v1, v5);
    Cx<V> t2 = cadd_t(v2, v6), u2 = csub_t(v2, v6);
    Cx<V> t3 = cadd_t(v3, v7), u3 = csub_t(v3, v7);
    u1 = {O::muls(O::add(u1.re, u1.im), h), O::muls(O::sub(u1.im, u1.re), h)};    // * (1 - i)/sqrt2
    u2 = cmul_mi_t(u2);                                                            // * -i
    u3 = {O::muls(O::sub(u3.im, u3.re), h), O::muls(O::add(u3.re, u3.im), -h)};   // * (-1 - i)/sqrt2
    dft4_t(t0, t1, t2, t3);
    dft4_t(u0, u1, u2, u3);
    v0 = t0; v2 = t1; v4 = t2; v6 = t3;
    v1 = u0; v3 = u1; v5 = u2; v7 = u3;
}

template <int J, typename V>
__device__ __forceinline__ Cx<V> mul_w32_t(Cx<V> a) {
    constexpr int j = J & 31;
    if constexpr (j == 0) {
        return a;
    } else if constexpr (j == 8) {
        return cmul_mi_t(a);
    } else if constexpr (j == 16) {
        return {Ops<V>::neg(a.re), Ops<V>::neg(a.im)};
    } else if constexpr (j == 24) {
        return {Ops<V>::neg(a.im), a.re};
    } else {
        return cmul_cs(a, kC32[j], -kS32[j]);
    }
}
template <int J, typename V>
__device__ __forceinline__ Cx<V> mul_w16_t(Cx<V> a) {
    constexpr int j = J & 15;
    if constexpr (j == 0) {
        return a;
    } else if constexpr (j == 4) {
        return cmul_mi_t(a);
    } else if constexpr (j == 8) {
        return {Ops<V>::neg(a.re), Ops<V>::neg(a.im)};
    } else if constexpr (j == 12) {
        return {Ops<V>::neg(a.im), a.re};
    } else {
        return cmul_cs(a, kC16[j], -kS16[j]);
    }
}

// 32-point DFT (radix 4 x 8), X[k] left in v[perm32(k)]
template <int B, typename V>
__device__ __forceinline__ void dft32_col_t(Cx<V> (&v)[32]) {
    dft4_t(v[B], v[B + 8], v[B + 16], v[B + 24]);
    v[B + 8] = mul_w32_t<B>(v[B + 8]);
    v[B + 16] = mul_w32_t<2 * B>(v[B + 16]);
    v[B + 24] = mul_w32_t<3 * B>(v[B + 24]);
}
template <typename V>
__device__ __forceinline__ void dft32_t(Cx<V> (&v)[32]) {
    dft32_col_t<0>(v); dft32_col_t<1>(v); dft32_col_t<2>(v); dft32_col_t<3>(v);
    dft32_col_t<4>(v); dft32_col_t<5>(v); dft32_col_t<6>(v); dft32_col_t<7>(v);
#pragma unroll
    for (int q = 0; q < 4; q++)
        dft8_t(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]);
}

// R1-point DFT of v[O .. O + R1): X[k] left in v[O + permr<R1>(k)]
template <int R1>
__device__ __forceinline__ constexpr int permr(int k) { return R1 == 16 ? 4 * (k & 3) + (k >> 2) : k; }
template <int O, int B, typename V>
__device__ __forceinline__ void dft16_col_t(Cx<V> (&v)[32]) {
    dft4_t(v[O + B], v[O + B + 4], v[O + B + 8], v[O + B + 12]);
    v[O + B + 4] = mul_w16_t<B>(v[O + B + 4]);
    v[O + B + 8] = mul_w16_t<2 * B>(v[O + B + 8]);
    v[O + B + 12] = mul_w16_t<3 * B>(v[O + B + 12]);
}
template <int R1, int O, typename V>
__device__ __forceinline__ void dft_r1(Cx<V> (&v)[32]) {
    if constexpr (R1 == 16) {
        dft16_col_t<O, 0>(v); dft16_col_t<O, 1>(v); dft16_col_t<O, 2>(v); dft16_col_t<O, 3>(v);
#pragma unroll
        for (int q = 0; q < 4; q++) dft4_t(v[O + 4 * q], v[O + 4 * q + 1], v[O + 4 * q + 2], v[O + 4 * q + 3]);
    } else {
        static_assert(R1 == 8, "R1 is 8 or 16");
        dft8_t(v[O], v[O + 1], v[O + 2], v[O + 3], v[O + 4], v[O + 5], v[O + 6], v[O + 7]);
    }
}
template <int R1, typename V>
__device__ __forceinline__ void dft_groups(Cx<V> (&v)[32]) {
    if constexpr (R1 == 16) {
        dft_r1<16, 0>(v);
        dft_r1<16, 16>(v);
    } else {
        dft_r1<8, 0>(v);
        dft_r1<8, 8>(v);
        dft_r1<8, 16>(v);
        dft_r1<8, 24>(v);
    }
}

// max over the R1 lanes of a group
template <int R1>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = R1 / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// element i of the tile as seen by the two variants: the packed kernel transposes the real plane, then the imaginary
// plane, through float2 (A, B) elements; the scalar kernel moves (re, im) in one go
struct WarpSmem {
    float *wpad;         // [n_fft]        0.5 * window, zero outside the taps
    float2 *tw1;         // [R1 - 1][32]   W_NC^(lane * k1), k1 = 1..R1-1
    float2 *tw2;         // [16][32]       W_n_fft^(k_own(j, lane))
    float2 *tiles;       // [NW][tile_f2]
    const uint32_t *ms;  // MelItems blob
};

// float2 slots of one warp's tile: the 32 x 33 transpose tile, later G magnitude arrays + the mel partial sums
template <typename V>
__host__ __device__ inline int warp_mag_stride(const PlanDev &p) {
    // V slots per group; a float array takes half a float2 slot per element
    const int n = (kMagBase + (p.n_mel ? p.mi_max_reach : p.nc) + 2 + 3) & ~3;
    return n;
}
template <typename V>
__host__ __device__ inline int warp_tile_f2(const PlanDev &p) {
    const int G = 2048 / p.n_fft;
    const int v_slots = G * warp_mag_stride<V>(p) + (p.n_mel ? 2 * p.mi_groups * 32 + 4 : 0);
    const int f2 = sizeof(V) == 8 ? v_slots : (v_slots + 1) / 2;
    const int t = f2 > 32 * kRow ? f2 : 32 * kRow;
    return (t + 3) & ~3;
}

// V = float2: NW warps, one persistent CTA per SM.  V = float: 8 warps, 2 CTAs per SM.
// LIST (scalar only): walk the rescue list the packed kernel left behind instead of the (descriptor, tile) grid.
// DIRECT (packed only): the plan's mel schedule is the band-major one -- a compile-time variant so that each packed kernel
// carries one schedule's code (instruction-cache footprint: DESIGN.md section 4); the scalar twin decides at run time.
template <typename V, int R1, bool MEL, bool I16, bool UNAL, bool LIST, int NW, bool DIRECT = false>
__global__ void __launch_bounds__(NW * 32, sizeof(V) == 8 ? 1 : 2)
    stft_warp_kernel(const PlanDev p, const TrackDesc *__restrict__ tracks, long long n_items, RescueList rescue) {
    using O = Ops<V>;
    constexpr bool kPacked = sizeof(V) == 8;
    constexpr int G = 32 / R1;              // groups per warp
    constexpr int FPG = O::kFrames;         // frames per group
    constexpr int FI = G * FPG;             // frames per warp step
    constexpr int NFFT = 64 * R1, NC = 32 * R1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (LIST) pdl_wait();   // launched under the tail of the packed kernel that fills the list

    const int tile_f2 = warp_tile_f2<V>(p);
    WarpSmem sm;
    sm.wpad = reinterpret_cast<float *>(smem_raw);
    sm.tw1 = reinterpret_cast<float2 *>(sm.wpad + NFFT);
    sm.tw2 = sm.tw1 + (R1 - 1) * 32;
    sm.tiles = sm.tw2 + 16 * 32;
    sm.ms = reinterpret_cast<const uint32_t *>(sm.tiles + NW * tile_f2);
    for (int i = threadIdx.x; i < NFFT / 4; i += blockDim.x) {
        float4 w = __ldg(reinterpret_cast<const float4 *>(p.fast_wpad) + i);
        if (kPacked && I16) {  // the 2^-15 of "s / 32768" rides on the window table (exact)
            const float k = 3.0517578125e-05f;
            w = make_float4(w.x * k, w.y * k, w.z * k, w.w * k);
        }
        reinterpret_cast<float4 *>(sm.wpad)[i] = w;
    }
    for (int i = threadIdx.x; i < ((R1 - 1) * 32 + 16 * 32) / 2; i += blockDim.x)
        reinterpret_cast<float4 *>(sm.tw1)[i] = __ldg(reinterpret_cast<const float4 *>(p.fast_tw) + i);
    for (int i = threadIdx.x; i < NW * tile_f2 / 2; i += blockDim.x)
        reinterpret_cast<float4 *>(sm.tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MEL) {
        for (int i = threadIdx.x; i < p.mi_words / 4; i += blockDim.x)
            reinterpret_cast<uint4 *>(const_cast<uint32_t *>(sm.ms))[i] = __ldg(reinterpret_cast<const uint4 *>(p.mi_blob) + i);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *tile = sm.tiles + warp * tile_f2;
    const int mag_stride = warp_mag_stride<V>(p);
    V *mag0 = reinterpret_cast<V *>(tile) + kMagBase;                 // group g: mag0 + g * mag_stride
    V *part = reinterpret_cast<V *>(tile) + G * mag_stride;
    const int half = p.win / 2;
    const int grp = lane / R1, k1 = lane % R1;
    const int partner = grp * R1 + ((R1 - k1) % R1);
    const int k1z = k1 ? k1 : R1;
    const long long tiles_per_track = rescue.tiles_per_track;
    const int tile_frames = static_cast<int>(rescue.tile_frames);
    if (LIST) n_items = min(*rescue.count, rescue.capacity);
    // rows of 64 FFT positions that hold a window tap: the reference zero-pads the windowed frame to n_fft (stft.rs:35-48),
    // so at 40 ms / 16 kHz (win 640 in n_fft 1024) six of the sixteen rows are zeros and are not loaded at all
    unsigned live_rows = 0;
#pragma unroll
    for (int n1 = 0; n1 < R1; n1++)
        if (p.load_all_rows || (64 * n1 + 63 >= p.pad_left && 64 * n1 < p.pad_left + p.win)) live_rows |= 1u << n1;

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        long long track, tile_idx;
        if (LIST) {
            const uint2 it = rescue.items[item];
            track = it.x;
            tile_idx = it.y;
        } else {
            track = item / tiles_per_track;
            tile_idx = item - track * tiles_per_track;
        }
        const TrackDesc d = tracks[track];
        const long long f_begin = tile_idx * tile_frames;
        if (f_begin >= d.n_frames) continue;
        const long long f_end = min(f_begin + tile_frames, d.n_frames);
        float lmax = -CUDART_INF_F, lnmin = -CUDART_INF_F;
        bool flagged = false;

        for (long long f0 = f_begin + static_cast<long long>(FI) * warp; f0 < f_end; f0 += static_cast<long long>(FI) * NW) {
            Cx<V> v[32];
            // ---- load + window: v[g R1 + n1] = z[32 n1 + lane] of group g's frame(s) ----
            // (a frame past f_end repeats the last one: its results are not stored)
            if constexpr (kPacked) {
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const long long fa = min(f0 + 2 * g, f_end - 1), fb = min(f0 + 2 * g + 1, f_end - 1);
                    const long long first_a = (d.frame_begin + fa) * p.hop - half - p.pad_left;  // file index of FFT position 0
                    const long long first_b = (d.frame_begin + fb) * p.hop - half - p.pad_left;
                    if constexpr (I16 && UNAL) {
                        const short *src_a = reinterpret_cast<const short *>(d.pcm) + (first_a - d.pcm_offset) + 2 * lane;
                        const short *src_b = reinterpret_cast<const short *>(d.pcm) + (first_b - d.pcm_offset) + 2 * lane;
#pragma unroll
                        for (int n1 = 0; n1 < R1; n1++) {
                            const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                            const float a_re = static_cast<float>(static_cast<int>(__ldg(src_a + 64 * n1)));
                            const float a_im = static_cast<float>(static_cast<int>(__ldg(src_a + 64 * n1 + 1)));
                            const float b_re = static_cast<float>(static_cast<int>(__ldg(src_b + 64 * n1)));
                            const float b_im = static_cast<float>(static_cast<int>(__ldg(src_b + 64 * n1 + 1)));
                            v[g * R1 + n1].re = make_float2(a_re * w.x, b_re * w.x);
                            v[g * R1 + n1].im = make_float2(a_im * w.y, b_im * w.y);
                        }
                    } else if constexpr (I16) {
                        const uint32_t *src_a = reinterpret_cast<const uint32_t *>(reinterpret_cast<const short *>(d.pcm) + (first_a - d.pcm_offset) + 2 * lane);
                        const uint32_t *src_b = reinterpret_cast<const uint32_t *>(reinterpret_cast<const short *>(d.pcm) + (first_b - d.pcm_offset) + 2 * lane);
#pragma unroll
                        for (int n1 = 0; n1 < R1; n1++) {
                            const uint32_t wa = __ldg(src_a + 32 * n1), wb = __ldg(src_b + 32 * n1);
                            const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                            const float a_re = static_cast<float>(static_cast<int>(static_cast<short>(wa & 0xffffu)));
                            const float a_im = static_cast<float>(static_cast<int>(wa) >> 16);
                            const float b_re = static_cast<float>(static_cast<int>(static_cast<short>(wb & 0xffffu)));
                            const float b_im = static_cast<float>(static_cast<int>(wb) >> 16);
                            v[g * R1 + n1].re = make_float2(a_re * w.x, b_re * w.x);
                            v[g * R1 + n1].im = make_float2(a_im * w.y, b_im * w.y);
                        }
                    } else if constexpr (UNAL) {
                        const float *src_a = d.pcm + (first_a - d.pcm_offset) + 2 * lane;
                        const float *src_b = d.pcm + (first_b - d.pcm_offset) + 2 * lane;
#pragma unroll
                        for (int n1 = 0; n1 < R1; n1++) {
                            const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                            v[g * R1 + n1].re = make_float2(__ldg(src_a + 64 * n1) * w.x, __ldg(src_b + 64 * n1) * w.x);
                            v[g * R1 + n1].im = make_float2(__ldg(src_a + 64 * n1 + 1) * w.y, __ldg(src_b + 64 * n1 + 1) * w.y);
                        }
                    } else {
                        const float *src_a = d.pcm + (first_a - d.pcm_offset) + 2 * lane;
                        const float *src_b = d.pcm + (first_b - d.pcm_offset) + 2 * lane;
#pragma unroll
                        for (int n1 = 0; n1 < R1; n1++) {
                            const bool live = (live_rows >> n1) & 1u;
                            const float2 zero2 = make_float2(0.0f, 0.0f);
                            const float2 xa = live ? __ldg(reinterpret_cast<const float2 *>(src_a + 64 * n1)) : zero2;
                            const float2 xb = live ? __ldg(reinterpret_cast<const float2 *>(src_b + 64 * n1)) : zero2;
                            const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                            v[g * R1 + n1].re = make_float2(xa.x * w.x, xb.x * w.x);
                            v[g * R1 + n1].im = make_float2(xa.y * w.y, xb.y * w.y);
                        }
                    }
                }
            } else {
                // scalar twin: any frame.  Interior frames take direct loads; file edges go through the warp's tile with
                // numpy-style reflect indices (utils.rs:111-137; taps outside the window are zero).
                float *stage = reinterpret_cast<float *>(tile);
                bool any_edge = false;
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const long long f = min(f0 + g, f_end - 1);
                    const long long tap0 = (d.frame_begin + f) * p.hop - half;
                    const long long first = tap0 - p.pad_left;
                    const bool interior = tap0 >= 0 && tap0 + p.win <= d.full_len && first >= d.pcm_offset &&
                                          first + NFFT <= d.pcm_offset + d.slice_len;
                    if (!interior) {
                        any_edge = true;
                        stage_edge_frame<NFFT>(d, p.pad_left, p.win, tap0, sm.wpad, stage + g * NFFT, lane);
                    }
                }
                if (any_edge) __syncwarp();
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const long long f = min(f0 + g, f_end - 1);
                    const long long tap0 = (d.frame_begin + f) * p.hop - half;
                    const long long first = tap0 - p.pad_left;
                    const bool interior = tap0 >= 0 && tap0 + p.win <= d.full_len && first >= d.pcm_offset &&
                                          first + NFFT <= d.pcm_offset + d.slice_len;
                    if (interior) {
                        const long long at = (first - d.pcm_offset) + 2 * lane;
#pragma unroll
                        for (int n1 = 0; n1 < R1; n1++) {
                            const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                            // the packed kernel's I16 path multiplies the INTEGER sample by (window * 2^-15); the exact
                            // power of two commutes with the rounding, so (s * 2^-15) * w is the same f32 value
                            const bool live = (live_rows >> n1) & 1u;
                            v[g * R1 + n1].re = (live ? pcm_sample(d, at + 64 * n1) : 0.0f) * w.x;
                            v[g * R1 + n1].im = (live ? pcm_sample(d, at + 64 * n1 + 1) : 0.0f) * w.y;
                        }
                    } else {
#pragma unroll
                        for (int n1 = 0; n1 < R1; n1++) {
                            const float2 x = *reinterpret_cast<const float2 *>(stage + g * NFFT + 64 * n1 + 2 * lane);
                            v[g * R1 + n1].re = x.x;
                            v[g * R1 + n1].im = x.y;
                        }
                    }
                }
                if (any_edge) __syncwarp();
            }
            // ---- pass 1: R1-point DFT of every group, twiddle W_NC^(lane k1), transpose ----
            dft_groups<R1>(v);
#pragma unroll
            for (int g = 0; g < G; g++) {
#pragma unroll
                for (int q = 1; q < R1; q++) {
                    const float2 w = sm.tw1[(q - 1) * 32 + lane];
                    v[g * R1 + permr<R1>(q)] = cmul_cs(v[g * R1 + permr<R1>(q)], w.x, w.y);
                }
            }
            if constexpr (kPacked) {
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int q = 0; q < R1; q++) tile[(g * R1 + q) * kRow + lane] = v[g * R1 + permr<R1>(q)].re;
                __syncwarp();
                Cx<V> t[32];
#pragma unroll
                for (int n2 = 0; n2 < 32; n2++) t[n2].re = tile[lane * kRow + n2];
                __syncwarp();
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int q = 0; q < R1; q++) tile[(g * R1 + q) * kRow + lane] = v[g * R1 + permr<R1>(q)].im;
                __syncwarp();
#pragma unroll
                for (int n2 = 0; n2 < 32; n2++) {
                    t[n2].im = tile[lane * kRow + n2];
                    v[n2] = t[n2];
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int q = 0; q < R1; q++)
                        tile[(g * R1 + q) * kRow + lane] = make_float2(v[g * R1 + permr<R1>(q)].re, v[g * R1 + permr<R1>(q)].im);
                __syncwarp();
#pragma unroll
                for (int n2 = 0; n2 < 32; n2++) {
                    const float2 x = tile[lane * kRow + n2];
                    v[n2].re = x.x;
                    v[n2].im = x.y;
                }
                __syncwarp();
            }
            // ---- pass 2: 32-point DFT over n2 -> lane (grp, k1) holds Z[k1 + R1 k2] in v[perm32(k2)] ----
            dft32_t(v);
            // ---- real split: 16 (k, NC - k) pairs per lane; |X| or dB ----
            // the frame(s) of this lane's group, and whether they exist
            const long long fr0 = f0 + static_cast<long long>(FPG) * grp;
            const bool ok0 = fr0 < f_end, ok1 = kPacked && fr0 + 1 < f_end;
            float *orow0 = d.out + min(fr0, f_end - 1) * p.n_bins;
            float *orow1 = d.out + min(fr0 + 1, f_end - 1) * p.n_bins;
            V *mag = mag0 + grp * mag_stride;
            float db_off = 0.0f;  // scalar twin: the dB shift of a rescaled frame
            bool group_ok = true;
            for (int attempt = 0;; attempt++) {
                V smax = O::zero();
                float fmx = -CUDART_INF_F, fnm = -CUDART_INF_F;  // this attempt's own max / -min (linear scale)
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const Cx<V> own_a = v[perm32(31 - j)], own_b = v[perm32((32 - j) & 31)];
                    Cx<V> zk, zn;
                    const Cx<V> sup = v[perm32(j)];
                    if constexpr (kPacked) {
                        zk.re.x = k1 ? own_a.re.x : own_b.re.x;
                        zk.re.y = k1 ? own_a.re.y : own_b.re.y;
                        zk.im.x = k1 ? own_a.im.x : own_b.im.x;
                        zk.im.y = k1 ? own_a.im.y : own_b.im.y;
                        zn.re.x = __shfl_sync(0xffffffffu, sup.re.x, partner);
                        zn.re.y = __shfl_sync(0xffffffffu, sup.re.y, partner);
                        zn.im.x = __shfl_sync(0xffffffffu, sup.im.x, partner);
                        zn.im.y = __shfl_sync(0xffffffffu, sup.im.y, partner);
                    } else {
                        zk.re = k1 ? own_a.re : own_b.re;
                        zk.im = k1 ? own_a.im : own_b.im;
                        zn.re = __shfl_sync(0xffffffffu, sup.re, partner);
                        zn.im = __shfl_sync(0xffffffffu, sup.im, partner);
                    }
                    const V er = O::add(zk.re, zn.re), ei = O::sub(zk.im, zn.im), dr = O::sub(zk.re, zn.re), di = O::add(zk.im, zn.im);
                    const float2 w = sm.tw2[j * 32 + lane];
                    const V wr = O::fmas(di, -w.y, O::muls(dr, w.x)), wi = O::fmas(dr, w.y, O::muls(di, w.x));
                    const V ar = O::add(er, wi), ai = O::sub(ei, wr), br = O::sub(er, wi), bi = O::add(ei, wr);
                    const V sa = O::fma(ar, ar, O::mul(ai, ai)), sb = O::fma(br, br, O::mul(bi, bi));
                    const int k_own = k1z + R1 * (31 - j), k_par = NC - k_own;
                    if constexpr (kPacked) {
                        smax.x = fmaxf(smax.x, fmaxf(sa.x, sb.x));
                        smax.y = fmaxf(smax.y, fmaxf(sa.y, sb.y));
                        if (MEL) {
                            mag[k_own] = make_float2(sqrt_ftz(sa.x), sqrt_ftz(sa.y));
                            mag[k_par] = make_float2(sqrt_ftz(sb.x), sqrt_ftz(sb.y));
                        } else {
                            const float2 a = O::muls(make_float2(lg2_ftz(sa.x), lg2_ftz(sa.y)), kDbPerLog2Pow);
                            const float2 b = O::muls(make_float2(lg2_ftz(sb.x), lg2_ftz(sb.y)), kDbPerLog2Pow);
                            if (ok0) {
                                orow0[k_own] = a.x;
                                orow0[k_par] = b.x;
                                fmx = fmaxf(fmx, fmaxf(a.x, b.x));
                                fnm = fmaxf(fnm, fmaxf(-a.x, -b.x));
                            }
                            if (ok1) {
                                orow1[k_own] = a.y;
                                orow1[k_par] = b.y;
                                fmx = fmaxf(fmx, fmaxf(a.y, b.y));
                                fnm = fmaxf(fnm, fmaxf(-a.y, -b.y));
                            }
                        }
                    } else {
                        smax = fmaxf(smax, fmaxf(sa, sb));
                        if (MEL) {
                            mag[k_own] = sqrt_ftz(sa);
                            mag[k_par] = sqrt_ftz(sb);
                        } else {
                            // the packed kernel rounds lg2 * c once (FMUL); an FMA with an offset of 0 rounds the same product once
                            const float a = fmaf(kDbPerLog2Pow, lg2_ftz(sa), db_off), b = fmaf(kDbPerLog2Pow, lg2_ftz(sb), db_off);
                            if (ok0) {
                                orow0[k_own] = a;
                                orow0[k_par] = b;
                                fmx = fmaxf(fmx, fmaxf(a, b));
                                fnm = fmaxf(fnm, fmaxf(-a, -b));
                            }
                        }
                    }
                }
                if (k1 == 0) {  // k = NC / 2 pairs with itself: X[NC/2] = conj(2 Z'[NC/2])
                    const Cx<V> z = v[perm32(16)];
                    const V s5 = O::muls(O::fma(z.re, z.re, O::mul(z.im, z.im)), 4.0f);
                    if constexpr (kPacked) {
                        smax.x = fmaxf(smax.x, s5.x);
                        smax.y = fmaxf(smax.y, s5.y);
                        if (MEL) {
                            mag[NC / 2] = make_float2(sqrt_ftz(s5.x), sqrt_ftz(s5.y));
                        } else {
                            const float2 a = O::muls(make_float2(lg2_ftz(s5.x), lg2_ftz(s5.y)), kDbPerLog2Pow);
                            if (ok0) {
                                orow0[NC / 2] = a.x;
                                fmx = fmaxf(fmx, a.x);
                                fnm = fmaxf(fnm, -a.x);
                            }
                            if (ok1) {
                                orow1[NC / 2] = a.y;
                                fmx = fmaxf(fmx, a.y);
                                fnm = fmaxf(fnm, -a.y);
                            }
                        }
                    } else {
                        smax = fmaxf(smax, s5);
                        if (MEL) {
                            mag[NC / 2] = sqrt_ftz(s5);
                        } else {
                            const float a = fmaf(kDbPerLog2Pow, lg2_ftz(s5), db_off);
                            if (ok0) {
                                orow0[NC / 2] = a;
                                fmx = fmaxf(fmx, a);
                                fnm = fmaxf(fnm, -a);
                            }
                        }
                    }
                }
                if constexpr (kPacked) {
                    // frames outside the exact range of the f32 square: the tile goes on the rescue list, and nothing of
                    // this group enters the min/max here (the scalar twin redoes the whole tile, bit-identically for
                    // the frames that were fine)
                    const float mx_a = group_max<R1>(smax.x), mx_b = group_max<R1>(smax.y);
                    group_ok = !((mx_a < kPowTiny && mx_a > 0.0f) || mx_a > kPowHuge || (mx_b < kPowTiny && mx_b > 0.0f) || mx_b > kPowHuge);
                    flagged |= !group_ok;
                    if (!MEL && group_ok) {
                        lmax = fmaxf(lmax, fmx);
                        lnmin = fmaxf(lnmin, fnm);
                    }
                    break;
                } else {
                    const float gm = group_max<R1>(smax);
                    const bool tiny = gm < kPowTiny && gm > 0.0f, huge = gm > kPowHuge;
                    const bool redo = attempt == 0 && (tiny || huge);
                    // every lane runs the same number of attempts (the shuffles above are warp-wide): a group that is
                    // fine repeats with scale 1 (exact), a group out of range with 2^(+-60) and the matching dB shift
                    if (!__any_sync(0xffffffffu, redo)) {
                        if (!MEL) {
                            lmax = fmaxf(lmax, fmx);
                            lnmin = fmaxf(lnmin, fnm);
                        }
                        break;
                    }
                    const float sc = redo ? (tiny ? kRescueUp : kRescueDown) : 1.0f;
                    if (redo) db_off = tiny ? -60.0f * kDbPerLog2Amp : 60.0f * kDbPerLog2Amp;
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        v[i].re *= sc;
                        v[i].im *= sc;
                    }
                }
            }
            if (MEL) {
                __syncwarp();
                const MelView mv(sm.ms);
#pragma unroll 1
                for (int g = 0; g < G; g++) {
                    const long long fg = f0 + static_cast<long long>(FPG) * g;
                    if (fg >= f_end) break;
                    if (kPacked ? !DIRECT : !mv.direct) {
                        mel_walk4<V>(mv, mag0 + g * mag_stride, part, lane);
                        __syncwarp();
                    }
                    float *row0 = d.out + fg * p.n_bins;
                    float *row1 = d.out + min(fg + 1, f_end - 1) * p.n_bins;
                    const bool has1 = kPacked && fg + 1 < f_end;
                    const bool g_ok = __shfl_sync(0xffffffffu, group_ok ? 1 : 0, g * R1) != 0;
                    const float g_off = __shfl_sync(0xffffffffu, db_off, g * R1);
                    auto emit = [&](int m, V acc) {
                        if (m >= mv.n_mel) return;
                        if constexpr (kPacked) {
                            const float2 db = O::muls(make_float2(lg2_ftz(acc.x), lg2_ftz(acc.y)), kDbPerLog2Amp);
                            row0[m] = db.x;
                            if (has1) row1[m] = db.y;
                            if (g_ok) {
                                lmax = fmaxf(lmax, has1 ? fmaxf(db.x, db.y) : db.x);
                                lnmin = fmaxf(lnmin, has1 ? fmaxf(-db.x, -db.y) : -db.x);
                            }
                        } else {
                            // exact zero stays -inf; the offset only shifts finite values (0 for an ordinary frame)
                            const float db = fmaf(kDbPerLog2Amp, lg2_ftz(acc), g_off);
                            row0[m] = db;
                            lmax = fmaxf(lmax, db);
                            lnmin = fmaxf(lnmin, -db);
                        }
                    };
                    if constexpr (kPacked && DIRECT) {
                        for (int r = 0; 32 * r < mv.n_mel; r += 2) {   // two rounds per walk (the host pads them in pairs)
                            V acc_a, acc_b;
                            mel_direct2<V>(mv, mag0 + g * mag_stride, r, lane, acc_a, acc_b);
                            emit(32 * r + lane, acc_a);
                            emit(32 * r + 32 + lane, acc_b);
                        }
                    } else if constexpr (kPacked) {
                        for (int r = 0; 32 * r < mv.n_mel; r++) emit(32 * r + lane, mel_band4<V>(mv, part, r, lane));
                    } else {
                        for (int r = 0; 32 * r < mv.n_mel; r++)
                            emit(32 * r + lane, mv.direct ? mel_direct<V>(mv, mag0 + g * mag_stride, r, lane) : mel_band4<V>(mv, part, r, lane));
                    }
                    __syncwarp();
                }
            }
        }
        // ---- per-channel {max, -min}: one atomic pair per warp and work item ----
        lmax = warp_max(lmax);
        lnmin = warp_max(lnmin);
        if (lane == 0) {
            if (lmax > -CUDART_INF_F || lnmin > -CUDART_INF_F) {
                atomic_max_float(&d.minmax[0], lmax);
                atomic_max_float(&d.minmax[1], lnmin);
            }
        }
        if constexpr (kPacked) {
            flagged = __any_sync(0xffffffffu, flagged);
            if (lane == 0 && flagged && atomicExch(&rescue.flags[item], 1u) == 0u) {
                const unsigned slot = atomicAdd(rescue.count, 1u);
                if (slot < rescue.capacity) rescue.items[slot] = make_uint2(static_cast<unsigned>(track), static_cast<unsigned>(tile_idx));
            }
        }
    }
}

// 10 packed warps: the default banks at 8 - 24 kHz carry up to 13 mel groups; with 12 warps tiles + tables pass 220 KB
constexpr int kPackedWarps = 10, kScalarWarps = 8;

template <typename V>
size_t warp_smem_bytes(const PlanDev &p, int nw) {
    const int r1 = p.n_fft / 64;
    return sizeof(float) * p.n_fft + sizeof(float2) * ((r1 - 1) * 32 + 16 * 32) + sizeof(float2) * nw * warp_tile_f2<V>(p) +
           sizeof(uint32_t) * static_cast<size_t>(p.n_mel ? p.mi_words : 0);
}

template <typename V, int R1, bool MEL, bool I16, bool UNAL, bool LIST, int NW, bool DIRECT = false>
cudaError_t launch_one(const PlanDev &plan, const TrackDesc *d_tracks, long long n_items, int grid, RescueList rescue, cudaStream_t st) {
    const size_t smem = warp_smem_bytes<V>(plan, NW);
    auto kern = stft_warp_kernel<V, R1, MEL, I16, UNAL, LIST, NW, DIRECT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    static const char *all_rows = getenv("THB_WARP_ALLROWS");
    PlanDev pl = plan;
    pl.load_all_rows = all_rows && atoi(all_rows) != 0;
    if (LIST) return launch_pdl(kern, dim3(grid), dim3(NW * 32), smem, st, pl, d_tracks, n_items, rescue);
    kern<<<grid, NW * 32, smem, st>>>(pl, d_tracks, n_items, rescue);
    return cudaGetLastError();
}

template <int R1, bool MEL, bool DIRECT>
cudaError_t launch_packed_md(const PlanDev &plan, const TrackDesc *d, long long n_items, int grid, RescueList rl, bool i16, bool unal, cudaStream_t st) {
    if (i16) return unal ? launch_one<float2, R1, MEL, true, true, false, kPackedWarps, DIRECT>(plan, d, n_items, grid, rl, st)
                         : launch_one<float2, R1, MEL, true, false, false, kPackedWarps, DIRECT>(plan, d, n_items, grid, rl, st);
    return unal ? launch_one<float2, R1, MEL, false, true, false, kPackedWarps, DIRECT>(plan, d, n_items, grid, rl, st)
                : launch_one<float2, R1, MEL, false, false, false, kPackedWarps, DIRECT>(plan, d, n_items, grid, rl, st);
}
template <int R1>
cudaError_t launch_packed_r1(const PlanDev &plan, const TrackDesc *d, long long n_items, int grid, RescueList rl, bool i16, bool unal, cudaStream_t st) {
    if (!plan.n_mel) return launch_packed_md<R1, false, false>(plan, d, n_items, grid, rl, i16, unal, st);
    return plan.mi_direct ? launch_packed_md<R1, true, true>(plan, d, n_items, grid, rl, i16, unal, st)
                          : launch_packed_md<R1, true, false>(plan, d, n_items, grid, rl, i16, unal, st);
}

template <int R1, bool LIST>
cudaError_t launch_scalar_r1(const PlanDev &plan, const TrackDesc *d, long long n_items, int grid, RescueList rl, cudaStream_t st) {
    if (plan.n_mel) return launch_one<float, R1, true, false, false, LIST, kScalarWarps>(plan, d, n_items, grid, rl, st);
    return launch_one<float, R1, false, false, false, LIST, kScalarWarps>(plan, d, n_items, grid, rl, st);
}

}  // namespace

// frames per work item of the packed kernel: four warp steps per warp
int stft_warp_tile_frames(const PlanDev &p) { return 4 * kPackedWarps * 2 * (2048 / p.n_fft); }

bool stft_warp_supported(const PlanDev &p) {
    if ((p.n_fft != 1024 && p.n_fft != 512) || !p.fast_wpad || !p.fast_tw) return false;
    if (p.n_mel && (!p.mi_blob || p.mi_min_start < -(k2048::kMagBase - 1))) return false;
    return warp_smem_bytes<float2>(p, kPackedWarps) <= 220 * 1024 && warp_smem_bytes<float>(p, kScalarWarps) <= 112 * 1024;
}

// the interior frames of every descriptor (whole n_fft span inside the slice, no reflect padding); `unaligned`: some
// frame does not start on a sample-pair boundary (odd hop / odd start): one load per sample
cudaError_t launch_stft_warp_packed(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, RescueList rescue, bool pcm_i16,
                                    bool unaligned, int sm_count, cudaStream_t st) {
    if (n_tracks <= 0 || rescue.tiles_per_track == 0) return cudaSuccess;
    const long long n_items = static_cast<long long>(n_tracks) * rescue.tiles_per_track;
    const int grid = static_cast<int>(n_items < sm_count ? n_items : sm_count);
    if (plan.n_fft == 1024) return launch_packed_r1<16>(plan, d_tracks, n_items, grid, rescue, pcm_i16, unaligned, st);
    return launch_packed_r1<8>(plan, d_tracks, n_items, grid, rescue, pcm_i16, unaligned, st);
}

// the scalar twin over whole descriptors (file edges, leftovers): tiles of rescue.tile_frames frames
cudaError_t launch_stft_warp_scalar(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, long long max_frames, int sm_count,
                                    cudaStream_t st) {
    if (n_tracks <= 0 || max_frames <= 0) return cudaSuccess;
    RescueList rl{};
    rl.tile_frames = 64;
    rl.tiles_per_track = static_cast<unsigned>((max_frames + 63) / 64);
    const long long n_items = static_cast<long long>(n_tracks) * rl.tiles_per_track;
    const int grid = static_cast<int>(n_items < 2ll * sm_count ? n_items : 2ll * sm_count);
    if (plan.n_fft == 1024) return launch_scalar_r1<16, false>(plan, d_tracks, n_items, grid, rl, st);
    return launch_scalar_r1<8, false>(plan, d_tracks, n_items, grid, rl, st);
}

// the scalar twin over the tiles on a rescue list (persistent grid; a no-op when the list is empty)
cudaError_t launch_stft_warp_list(const PlanDev &plan, const TrackDesc *d_tracks, RescueList rescue, int sm_count, cudaStream_t st) {
    if (plan.n_fft == 1024) return launch_scalar_r1<16, true>(plan, d_tracks, 0, 2 * sm_count, rescue, st);
    return launch_scalar_r1<8, true>(plan, d_tracks, 0, 2 * sm_count, rescue, st);
}

}  // namespace thb
