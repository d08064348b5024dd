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
    const uint2 *grp;      // {T, word offset of the group's weights}
    const int32_t *start;  // [n_groups * 32]
    const uint2 *rounds;   // {K, first row} per 32 bands
    const uint16_t *goff;  // rows of 32 slot ids
    const uint2 *rounds4;  // {rows of four, first row of four} per 32 bands
    const uint4 *goff4;    // rows of 32 lanes x 4 byte offsets (float2 slots)
    int direct;            // the band-major schedule is the cheaper one for this bank: mel_direct instead of walk + gather
    const uint2 *drounds;  // {steps, word offset of the round's weights} per 32 bands
    const int32_t *dk0;    // [rounds * 32] first bin of the lane's band
    __device__ __forceinline__ explicit MelView(const uint32_t *b) : base(b) {
        n_groups = static_cast<int>(b[0]);
        n_mel = static_cast<int>(b[1]);
        grp = reinterpret_cast<const uint2 *>(b + b[2]);
        start = reinterpret_cast<const int32_t *>(b + b[3]);
        rounds = reinterpret_cast<const uint2 *>(b + b[4]);
        goff = reinterpret_cast<const uint16_t *>(b + b[5]);
        rounds4 = reinterpret_cast<const uint2 *>(b + b[8]);
        goff4 = reinterpret_cast<const uint4 *>(b + b[9]);
        direct = static_cast<int>(b[10]);
        drounds = reinterpret_cast<const uint2 *>(b + b[11]);
        dk0 = reinterpret_cast<const int32_t *>(b + b[12]);
    }
};

// elements (float for the scalar kernel, float2 for the pair kernel) of one warp's tile: the transpose tile,
// later the magnitudes [kMagBase + bin] with the mel walk's lead / reach, then two partial sums (rise, fall) per
// mel slot: part[slot] and part[n_slots + slot]
__host__ __device__ inline int part_base(const PlanDev &p) { return (kMagBase + (p.n_mel ? p.mi_max_reach : 1024) + 2) & ~1; }
__host__ __device__ inline int tile_elems(const PlanDev &p) {
    const int need = part_base(p) + (p.n_mel ? 2 * p.mi_groups * 32 + 1 : 0);  // + the always-zero slot
    const int t = need > 32 * kRow ? need : 32 * kRow;
    return (t + 3) & ~3;
}

// The sparse mel product of one frame (V = float) or one frame pair (V = float2) from the warp's magnitudes,
// following the MelItems schedule (thb_host.hpp).  Leaves the band sums in part[...]: the caller gathers them.
template <typename V>
struct MelOps;
template <>
struct MelOps<float> {
    static __device__ __forceinline__ float zero() { return 0.0f; }
    static __device__ __forceinline__ float fma(float m, float w, float acc) { return fmaf(m, w, acc); }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
};
template <>
struct MelOps<float2> {
    static __device__ __forceinline__ float2 zero() { return make_float2(0.0f, 0.0f); }
    static __device__ __forceinline__ float2 fma(float2 m, float w, float2 acc) { return __ffma2_rn(m, make_float2(w, w), acc); }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
};

// (g0, gs): this warp takes groups g0, g0 + gs, ... -- (0, 1) when one warp owns the whole frame
template <typename V>
__device__ __forceinline__ void mel_walk(const MelView &mv, const V *mag, V *part, int lane, int g0 = 0, int gs = 1) {
    using O = MelOps<V>;
    const int n_slots = mv.n_groups * 32;
    for (int g = g0; g < mv.n_groups; g += gs) {
        const uint2 gh = mv.grp[g];
        const float4 *wq = reinterpret_cast<const float4 *>(mv.base + gh.y) + lane;
        const V *mq = mag + mv.start[g * 32 + lane];
        const int T2 = static_cast<int>(gh.x) >> 1;
        V rise = O::zero(), fall = O::zero();
#pragma unroll 2
        for (int t = 0; t < T2; t++) {
            const float4 w = wq[32 * t];
            const V m0 = mq[2 * t], m1 = mq[2 * t + 1];
            rise = O::fma(m0, w.x, rise);
            fall = O::fma(m0, w.y, fall);
            rise = O::fma(m1, w.z, rise);
            fall = O::fma(m1, w.w, fall);
        }
        part[g * 32 + lane] = rise;
        part[n_slots + g * 32 + lane] = fall;
    }
    if (lane == 0 && g0 == 0) part[2 * n_slots] = O::zero();  // the padding slot of the gather (the tile is reused by the transposes)
}

// band 32 r + lane: sum of its partial sums in the schedule's order (the padding rows add the always-zero slot)
template <typename V>
__device__ __forceinline__ V mel_band(const MelView &mv, const V *part, int r, int lane) {
    using O = MelOps<V>;
    const uint2 rd = mv.rounds[r];
    const uint16_t *row = mv.goff + rd.y * 32 + lane;
    V acc = O::zero();
    for (uint32_t j = 0; j < rd.x; j++) acc = O::add(acc, part[row[32 * j]]);
    return acc;
}

// ---- the same walk and gather with leaner bookkeeping: pointers advance instead of being recomputed, four steps per
// trip plus one optional two-step tail, four gather entries per 16-byte load of byte offsets.  The operations and their
// order are those of mel_walk / mel_band (padding entries add the always-zero slot), so the results are the same bit
// for bit.
template <typename V>
__device__ __forceinline__ void mel_walk4(const MelView &mv, const V *mag, V *part, int lane) {
    using O = MelOps<V>;
    const int n_slots = mv.n_groups * 32;
    const float4 *wlane = reinterpret_cast<const float4 *>(mv.base) + lane;
    V *prow = part + lane;
#pragma unroll 1
    for (int g = 0; g < mv.n_groups; g++) {
        const uint2 gh = mv.grp[g];
        const float4 *wq = wlane + (gh.y >> 2);
        const V *mq = mag + mv.start[g * 32 + lane];
        const int T4 = static_cast<int>(gh.x) >> 2;
        V rise = O::zero(), fall = O::zero();
#pragma unroll 1
        for (int t = 0; t < T4; t++) {
            const float4 w0 = wq[0], w1 = wq[32];
            const V m0 = mq[0], m1 = mq[1], m2 = mq[2], m3 = mq[3];
            wq += 64;
            mq += 4;
            rise = O::fma(m0, w0.x, rise);
            fall = O::fma(m0, w0.y, fall);
            rise = O::fma(m1, w0.z, rise);
            fall = O::fma(m1, w0.w, fall);
            rise = O::fma(m2, w1.x, rise);
            fall = O::fma(m2, w1.y, fall);
            rise = O::fma(m3, w1.z, rise);
            fall = O::fma(m3, w1.w, fall);
        }
        if (gh.x & 2u) {  // step counts are even
            const float4 w0 = wq[0];
            const V m0 = mq[0], m1 = mq[1];
            rise = O::fma(m0, w0.x, rise);
            fall = O::fma(m0, w0.y, fall);
            rise = O::fma(m1, w0.z, rise);
            fall = O::fma(m1, w0.w, fall);
        }
        prow[0] = rise;
        prow[n_slots] = fall;
        prow += 32;
    }
    if (lane == 0) part[2 * n_slots] = O::zero();  // the padding slot of the gather (the tile is reused by the transposes)
}

template <typename V>
__device__ __forceinline__ V mel_band4(const MelView &mv, const V *part, int r, int lane) {
    using O = MelOps<V>;
    constexpr int kShift = sizeof(V) == 8 ? 0 : 1;  // the table holds float2 byte offsets
    const uint2 rd = mv.rounds4[r];
    const uint4 *row = mv.goff4 + rd.y * 32 + lane;
    const unsigned char *pb = reinterpret_cast<const unsigned char *>(part);
    V acc = O::zero();
#pragma unroll 1
    for (uint32_t j = 0; j < rd.x; j++) {
        const uint4 o = row[32 * j];
        acc = O::add(acc, *reinterpret_cast<const V *>(pb + (o.x >> kShift)));
        acc = O::add(acc, *reinterpret_cast<const V *>(pb + (o.y >> kShift)));
        acc = O::add(acc, *reinterpret_cast<const V *>(pb + (o.z >> kShift)));
        acc = O::add(acc, *reinterpret_cast<const V *>(pb + (o.w >> kShift)));
    }
    return acc;
}

// Band-major mel product for banks of narrow bands (MelItems::use_direct): lane l owns band 32 r + l and walks its own
// bins, rd.x steps (a multiple of four: the round's longest band; shorter bands run on zero weights).  Four steps per
// trip into two accumulators (even / odd bins), all eight loads ahead of the first use: a single dependent chain of
// load -> FMA per step measured 8.6 instruction slots per step (latency bound).
template <typename V>
__device__ __forceinline__ V mel_direct(const MelView &mv, const V *mag, int r, int lane) {
    using O = MelOps<V>;
    const uint2 rd = mv.drounds[r];
    const float *w = reinterpret_cast<const float *>(mv.base + rd.y) + lane;
    const V *mq = mag + mv.dk0[32 * r + lane];
    V acc0 = O::zero(), acc1 = O::zero();
#pragma unroll 1
    for (uint32_t i = 0; i < rd.x; i += 4) {
        const float w0 = w[0], w1 = w[32], w2 = w[64], w3 = w[96];
        const V m0 = mq[0], m1 = mq[1], m2 = mq[2], m3 = mq[3];
        w += 128;
        mq += 4;
        acc0 = O::fma(m0, w0, acc0);
        acc1 = O::fma(m1, w1, acc1);
        acc0 = O::fma(m2, w2, acc0);
        acc1 = O::fma(m3, w3, acc1);
    }
    return O::add(acc0, acc1);
}

// Two rounds (bands 32 r + l and 32 r + 32 + l, r even) in one walk, for schedules built with direct_pairs (both rounds
// of a pair have the same step count): one loop carries two independent load -> FMA chains per lane (four
// accumulators).  Each band's sum is formed exactly as in mel_direct -- same steps, same order, same two accumulators --
// so the scalar kernel, which still goes round by round, agrees bit for bit.  One round at a time measured ~390 cycles
// per round and warp at 10 resident warps: the latency of table word -> addresses -> eight loads -> FMA chain -> MUFU ->
// store with nothing to overlap it with.  Measured (ms, 32 ch x 10 min, default bank): 16 kHz 3.99 -> 3.58, 8 kHz
// 2.40 -> 2.19, 22.05 kHz 3.77 -> 3.41, 24 kHz 3.68 -> 3.32; without the padding (joint loop, then the longer round alone)
// 3.78 / 2.32 / 3.61 / 3.50 -- the lone tail is the slow part.
template <typename V>
__device__ __forceinline__ void mel_direct2(const MelView &mv, const V *mag, int r, int lane, V &acc_a, V &acc_b) {
    using O = MelOps<V>;
    const uint4 rd = *reinterpret_cast<const uint4 *>(mv.drounds + r);  // {steps, weights of r, steps, weights of r + 1}
    const float *wa = reinterpret_cast<const float *>(mv.base + rd.y) + lane;
    const float *wb = reinterpret_cast<const float *>(mv.base + rd.w) + lane;
    const V *ma = mag + mv.dk0[32 * r + lane];
    const V *mb = mag + mv.dk0[32 * r + 32 + lane];
    V a0 = O::zero(), a1 = O::zero(), b0 = O::zero(), b1 = O::zero();
#pragma unroll 1
    for (uint32_t i = 0; i < rd.x; i += 4) {
        const float wa0 = wa[0], wa1 = wa[32], wa2 = wa[64], wa3 = wa[96];
        const float wb0 = wb[0], wb1 = wb[32], wb2 = wb[64], wb3 = wb[96];
        const V ma0 = ma[0], ma1 = ma[1], ma2 = ma[2], ma3 = ma[3];
        const V mb0 = mb[0], mb1 = mb[1], mb2 = mb[2], mb3 = mb[3];
        wa += 128;
        wb += 128;
        ma += 4;
        mb += 4;
        a0 = O::fma(ma0, wa0, a0);
        b0 = O::fma(mb0, wb0, b0);
        a1 = O::fma(ma1, wa1, a1);
        b1 = O::fma(mb1, wb1, b1);
        a0 = O::fma(ma2, wa2, a0);
        b0 = O::fma(mb2, wb2, b0);
        a1 = O::fma(ma3, wa3, a1);
        b1 = O::fma(mb3, wb3, b1);
    }
    acc_a = O::add(a0, a1);
    acc_b = O::add(b0, b1);
}

}  // namespace k2048

// A file-edge frame of the warp-register kernels: stage[pos] = sample(reflect(tap0 + pos - pad_left)) * wpad[pos] for the FFT
// positions pos = lane, lane + 32, ... < NFFT that hold a window tap, 0 for the others (numpy-style reflect, utils.rs:111-137;
// the reference zero-pads the windowed frame, stft.rs:35-48).  Sixteen positions per lane at a time -- indices, then
// the loads, then the products -- so that sixteen loads are in flight: one position after the other is one DRAM round
// trip each on samples no one has touched yet (64 x ~0.5 us = the 31 - 35 us an edge launch used to take).
template <int NFFT>
__device__ __forceinline__ void stage_edge_frame(const TrackDesc &d, int pad_left, int win, long long tap0, const float *wpad,
                                                 float *stage, int lane) {
    constexpr int kPerLane = NFFT / 32, kBatch = kPerLane < 16 ? kPerLane : 16;
    for (int b0 = 0; b0 < kPerLane; b0 += kBatch) {
        long long idx[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; j++) {
            const int a = (b0 + j) * 32 + lane - pad_left;
            const int ac = a < 0 ? 0 : (a >= win ? win - 1 : a);   // a tap of the window in every case: a valid address
            idx[j] = reflect_index(tap0 + ac, d.full_len) - d.pcm_offset;
        }
        float x[kBatch];
        if (d.pcm_i16) {
#pragma unroll
            for (int j = 0; j < kBatch; j++)
                x[j] = static_cast<float>(__ldg(reinterpret_cast<const short *>(d.pcm) + idx[j])) * 3.0517578125e-05f;
        } else {
#pragma unroll
            for (int j = 0; j < kBatch; j++) x[j] = __ldg(d.pcm + idx[j]);
        }
#pragma unroll
        for (int j = 0; j < kBatch; j++) {
            const int pos = (b0 + j) * 32 + lane, a = pos - pad_left;
            stage[pos] = (a >= 0 && a < win) ? x[j] * wpad[pos] : 0.0f;
        }
    }
}

}  // namespace thb
