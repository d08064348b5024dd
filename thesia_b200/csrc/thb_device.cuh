// thb_device.cuh -- small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdint>

namespace thb {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i:  (x + iy)(-i) = y - ix
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// Forward DFTs on registers, natural-order output: v[q] = sum_r v[r] * exp(-2 pi i q r / R)
__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
    const float2 s13 = cadd(a1, a3), d13 = cmul_mi(csub(a1, a3));
    a0 = cadd(s02, s13);
    a1 = cadd(d02, d13);
    a2 = csub(s02, s13);
    a3 = csub(d02, d13);
}
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    // decimation in frequency: t_r = v_r + v_{r+4}, u_r = (v_r - v_{r+4}) * w8^r;
    // even outputs = DFT4(t), odd outputs = DFT4(u)
    const float h = 0.70710678118654752440f;
    float2 t0 = cadd(v[0], v[4]), u0 = csub(v[0], v[4]);
    float2 t1 = cadd(v[1], v[5]), u1 = csub(v[1], v[5]);
    float2 t2 = cadd(v[2], v[6]), u2 = csub(v[2], v[6]);
    float2 t3 = cadd(v[3], v[7]), u3 = csub(v[3], v[7]);
    u1 = make_float2(h * (u1.x + u1.y), h * (u1.y - u1.x));    // * (1 - i)/sqrt2
    u2 = cmul_mi(u2);                                           // * -i
    u3 = make_float2(h * (u3.y - u3.x), -h * (u3.x + u3.y));   // * (-1 - i)/sqrt2
    dft4(t0, t1, t2, t3);
    dft4(u0, u1, u2, u3);
    v[0] = t0; v[2] = t1; v[4] = t2; v[6] = t3;
    v[1] = u0; v[3] = u1; v[5] = u2; v[7] = u3;
}

// numpy-style 'reflect' index (edge sample not repeated, period 2n-2), any s, n >= 2
// = Pad::pad(.., PadMode::Reflect) of utils.rs:111-137 including its multi-wrap cycle().
__device__ __forceinline__ long long reflect_index(long long s, long long n) {
    if (s >= 0 && s < n) return s;
    // one reflection (every frame of a file at least half a window long): no 64-bit division
    if (s < 0 && -s < n) return -s;
    if (s >= n && s <= 2 * (n - 1)) return 2 * (n - 1) - s;
    const long long period = 2 * (n - 1);
    long long m = s % period;
    if (m < 0) m += period;
    return m < n ? m : period - m;
}

// |re + i im| = Complex::norm (hypot); the rescaled leg keeps tiny/huge inputs exact enough.
__device__ __forceinline__ float cabs_safe(float re, float im) {
    const float s = fmaf(re, re, im * im);
    if (s > 1e-30f && s < 1e30f) return sqrtf(s);
    return hypotf(re, im);
}

// dB_from_amp_inplace_default for x >= 0: log10(x) * 20; 0 -> -inf; NaN -> NaN (decibel.rs:176-202)
__device__ __forceinline__ float amp_to_db(float x) { return log10f(x) * 20.0f; }

// float atomic max that also orders negatives and -inf (the slot starts at -inf)
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v != v) return;
    if (!signbit(v)) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// Programmatic dependent launch (sm_90+).  A kernel launched through launch_pdl() (thb_kernels.cuh) may be scheduled while
// the kernel before it in the stream is still draining; pdl_wait() returns once that kernel has completed and its writes
// are visible -- it is the first statement of every kernel launched that way, and a no-op in a kernel launched the
// ordinary way.  pdl_launch_dependents() lets the NEXT kernel's CTAs become resident (they then sit in their pdl_wait).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace thb
