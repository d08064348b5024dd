// thb_packed.cuh -- packed f32x2 complex arithmetic shared by the frame-pair kernels (thb_stft_pair.cu: n_fft 2048,
// thb_stft_big.cu: n_fft 16384): every register pair holds the same quantity of two consecutive frames
// (.x = frame A, .y = frame B) and all arithmetic is issued as sm_100 FADD2 / FMUL2 / FFMA2.
#pragma once
#include "thb_stft2048.cuh"

namespace thb {
namespace packed {

using namespace k2048;

// ---- packed helpers: .x = frame A, .y = frame B ------------------------------------------------
using f2 = float2;
__device__ __forceinline__ f2 neg(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 padd(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 psub(f2 a, f2 b) { return __fadd2_rn(a, neg(b)); }
__device__ __forceinline__ f2 pmul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 pfma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 bc(float s) { return make_float2(s, s); }  // broadcast operand

struct cx {
    f2 re, im;
};
__device__ __forceinline__ cx cadd2(cx a, cx b) { return {padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ cx csub2(cx a, cx b) { return {psub(a.re, b.re), psub(a.im, b.im)}; }
__device__ __forceinline__ cx cmul_mi2(cx a) { return {a.im, neg(a.re)}; }  // * -i
// a * (c + i s) with scalar c, s shared by both frames
__device__ __forceinline__ cx cmul_s(cx a, float c, float s) {
    cx r;
    r.re = pfma(a.im, bc(-s), pmul(a.re, bc(c)));
    r.im = pfma(a.re, bc(s), pmul(a.im, bc(c)));
    return r;
}

// a * W_32^J,  W_32 = exp(-2 pi i / 32) = cos - i sin
template <int J>
__device__ __forceinline__ cx mul_w32(cx a) {
    constexpr int j = J & 31;
    if constexpr (j == 0) {
        return a;
    } else if constexpr (j == 8) {
        return cmul_mi2(a);
    } else if constexpr (j == 16) {
        return {neg(a.re), neg(a.im)};
    } else if constexpr (j == 24) {
        return {neg(a.im), a.re};
    } else {
        return cmul_s(a, kC32[j], -kS32[j]);
    }
}

__device__ __forceinline__ void dft4p(cx &a0, cx &a1, cx &a2, cx &a3) {
    const cx s02 = cadd2(a0, a2), d02 = csub2(a0, a2);
    const cx s13 = cadd2(a1, a3), d13 = cmul_mi2(csub2(a1, a3));
    a0 = cadd2(s02, s13);
    a1 = cadd2(d02, d13);
    a2 = csub2(s02, s13);
    a3 = csub2(d02, d13);
}

__device__ __forceinline__ void dft8p(cx &v0, cx &v1, cx &v2, cx &v3, cx &v4, cx &v5, cx &v6, cx &v7) {
    const float h = 0.70710678118654752440f;
    cx t0 = cadd2(v0, v4), u0 = csub2(v0, v4);
    cx t1 = cadd2(v1, v5), u1 = csub2(v1, v5);
    cx t2 = cadd2(v2, v6), u2 = csub2(v2, v6);
    cx t3 = cadd2(v3, v7), u3 = csub2(v3, v7);
    u1 = {pmul(padd(u1.re, u1.im), bc(h)), pmul(psub(u1.im, u1.re), bc(h))};    // * (1 - i)/sqrt2
    u2 = cmul_mi2(u2);                                                            // * -i
    u3 = {pmul(psub(u3.im, u3.re), bc(h)), pmul(padd(u3.re, u3.im), bc(-h))};   // * (-1 - i)/sqrt2
    dft4p(t0, t1, t2, t3);
    dft4p(u0, u1, u2, u3);
    v0 = t0; v2 = t1; v4 = t2; v6 = t3;
    v1 = u0; v3 = u1; v5 = u2; v7 = u3;
}

// In-register 32-point forward DFT of both frames.  Output X[k] is left in v[perm32(k)].
template <int B>
__device__ __forceinline__ void dft32_col(cx (&v)[32]) {
    dft4p(v[B], v[B + 8], v[B + 16], v[B + 24]);
    v[B + 8] = mul_w32<B>(v[B + 8]);
    v[B + 16] = mul_w32<2 * B>(v[B + 16]);
    v[B + 24] = mul_w32<3 * B>(v[B + 24]);
}

__device__ __forceinline__ void dft32p(cx (&v)[32]) {
    // n = b + 8a, k = q + 4r:  W32^(nk) = W4^(aq) W32^(bq) W8^(br)
    dft32_col<0>(v); dft32_col<1>(v); dft32_col<2>(v); dft32_col<3>(v);
    dft32_col<4>(v); dft32_col<5>(v); dft32_col<6>(v); dft32_col<7>(v);
#pragma unroll
    for (int q = 0; q < 4; q++)
        dft8p(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6],
              v[8 * q + 7]);
}


// cos / sin of 2 pi j / 16
__device__ constexpr float kC16[16] = {1.0f, 0.9238795042037964f, 0.7071067690849304f, 0.3826834261417389f, 0.0f,
                                       -0.3826834261417389f, -0.7071067690849304f, -0.9238795042037964f, -1.0f,
                                       -0.9238795042037964f, -0.7071067690849304f, -0.3826834261417389f, 0.0f,
                                       0.3826834261417389f, 0.7071067690849304f, 0.9238795042037964f};
__device__ constexpr float kS16[16] = {0.0f, 0.3826834261417389f, 0.7071067690849304f, 0.9238795042037964f, 1.0f,
                                       0.9238795042037964f, 0.7071067690849304f, 0.3826834261417389f, 0.0f,
                                       -0.3826834261417389f, -0.7071067690849304f, -0.9238795042037964f, -1.0f,
                                       -0.9238795042037964f, -0.7071067690849304f, -0.3826834261417389f};

// a * W_16^J
template <int J>
__device__ __forceinline__ cx mul_w16(cx a) {
    constexpr int j = J & 15;
    if constexpr (j == 0) {
        return a;
    } else if constexpr (j == 4) {
        return cmul_mi2(a);
    } else if constexpr (j == 8) {
        return {neg(a.re), neg(a.im)};
    } else if constexpr (j == 12) {
        return {neg(a.im), a.re};
    } else {
        return cmul_s(a, kC16[j], -kS16[j]);
    }
}

// position of output X[k] of dft16p
__device__ __forceinline__ constexpr int perm16(int k) { return 4 * (k & 3) + (k >> 2); }

template <int B>
__device__ __forceinline__ void dft16_col(cx (&v)[16]) {
    dft4p(v[B], v[B + 4], v[B + 8], v[B + 12]);
    v[B + 4] = mul_w16<B>(v[B + 4]);
    v[B + 8] = mul_w16<2 * B>(v[B + 8]);
    v[B + 12] = mul_w16<3 * B>(v[B + 12]);
}

// In-register 16-point forward DFT of both frames (radix 4 x 4).  Output X[k] is left in v[perm16(k)].
__device__ __forceinline__ void dft16p(cx (&v)[16]) {
    // n = b + 4a, k = q + 4r:  W16^(nk) = W4^(aq) W16^(bq) W4^(br)
    dft16_col<0>(v); dft16_col<1>(v); dft16_col<2>(v); dft16_col<3>(v);
#pragma unroll
    for (int q = 0; q < 4; q++) dft4p(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

}  // namespace packed
}  // namespace thb
