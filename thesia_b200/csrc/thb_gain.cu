// thb_gain.cu -- gain normalisation + guard clipping ahead of the STFT:
// AudioTrack::apply_gain (track.rs:152-171: y = gain * x over the ORIGINAL samples) followed by the guard clipping of
// Audio::mutate (audio.rs:49-63) in its two elementwise modes, Clip (audio.rs:134-144) and ReduceGlobalLevel
// (audio.rs:146-160), and by the statistics `mutate` recomputes (dynamics/stats.rs:56-85, level part).
//
// HBM-bound: 4 bytes read + 4 written per sample (+ 4 for the optional WavBeforeClip copy); everything else rides
// along in the same pass:
//   pass P (ReduceGlobalLevel only)  peak of |gain * x| per track group          -- read only
//   pass A                           y = gain * x -> guard -> out, GuardClippingStats, sum of squares / abs max
// max and counts are exact and order independent (integer atomics on the bit patterns of non-negative floats);
// the f64 sums of squares go through per-CTA partials that a one-warp-per-channel kernel adds in a fixed order.
#include "thb_device.cuh"
#include "thb_kernels.cuh"

namespace thb {
namespace {

constexpr int kGainThreads = 256;
constexpr long long kGainChunk = 1 << 17;  // samples per CTA

__device__ __forceinline__ float clamp_unit(float x) {  // f32::clamp(-1, 1): NaN stays NaN
    return x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Acc {
    double ss = 0.0;           // sum of squares of the output
    float mx_out = 0.0f;       // abs max of the output
    float mx_before = 0.0f;    // abs max of gain * x
    unsigned cnt = 0;          // samples with |gain * x| > 1
};

// MODE 0: Clip.  1: ReduceGlobalLevel (peak known).  2: copy (unit / non-finite gain).  3: peak pass (read only).
template <int MODE>
__device__ __forceinline__ float gain_one(float x, float gain, double reduce, bool reducing, Acc &a, float &before) {
    if (MODE == 2) {
        a.ss += static_cast<double>(x) * x;
        a.mx_out = fmaxf(a.mx_out, fabsf(x));
        return x;
    }
    const float y = __fmul_rn(gain, x);
    before = y;
    if (MODE == 3) {
        a.mx_before = fmaxf(a.mx_before, fabsf(y));
        return y;
    }
    float o;
    if (MODE == 0) {
        a.mx_before = fmaxf(a.mx_before, fabsf(y));
        a.cnt += fabsf(y) > 1.0f;
        o = clamp_unit(y);
    } else {
        o = reducing ? clamp_unit(static_cast<float>(static_cast<double>(y) * reduce)) : y;
    }
    a.ss += static_cast<double>(o) * o;
    a.mx_out = fmaxf(a.mx_out, fabsf(o));
    return o;
}

template <int MODE>
__global__ void __launch_bounds__(kGainThreads) gain_kernel(const GainDesc *__restrict__ descs, int chunks_per_ch,
                                                            const unsigned *__restrict__ group_peak_bits,
                                                            double *__restrict__ part_ss, GainOut *__restrict__ outs,
                                                            unsigned *__restrict__ group_peak_out) {
    __shared__ double red_s[kGainThreads / 32];
    const GainDesc d = descs[blockIdx.y];
    const long long lo = static_cast<long long>(blockIdx.x) * kGainChunk;
    const long long hi = min(lo + kGainChunk, d.len);
    Acc a;
    double reduce = 1.0;
    bool reducing = false;
    if (MODE == 1) {
        const float peak = __uint_as_float(group_peak_bits[d.group]);
        reducing = static_cast<double>(peak) > 1.0;   // audio.rs:147-148
        if (reducing) reduce = 1.0 / static_cast<double>(peak);
    }
    if (lo < hi) {
        const bool vec = !d.pcm_i16 && ((reinterpret_cast<uintptr_t>(d.in) | reinterpret_cast<uintptr_t>(d.out) |
                                         reinterpret_cast<uintptr_t>(d.before)) & 15) == 0;
        if (vec) {
            const float4 *p = reinterpret_cast<const float4 *>(static_cast<const float *>(d.in) + lo);
            float4 *q = MODE == 3 ? nullptr : reinterpret_cast<float4 *>(d.out + lo);
            float4 *b = (MODE == 0 && d.before) ? reinterpret_cast<float4 *>(d.before + lo) : nullptr;
            const long long n4 = (hi - lo) >> 2;
            for (long long i = threadIdx.x; i < n4; i += kGainThreads) {
                const float4 v = __ldg(p + i);
                float4 o, y;
                o.x = gain_one<MODE>(v.x, d.gain, reduce, reducing, a, y.x);
                o.y = gain_one<MODE>(v.y, d.gain, reduce, reducing, a, y.y);
                o.z = gain_one<MODE>(v.z, d.gain, reduce, reducing, a, y.z);
                o.w = gain_one<MODE>(v.w, d.gain, reduce, reducing, a, y.w);
                if (MODE != 3) q[i] = o;
                if (MODE == 0 && b) b[i] = y;
            }
            for (long long i = lo + 4 * n4 + threadIdx.x; i < hi; i += kGainThreads) {
                float y;
                const float o = gain_one<MODE>(__ldg(static_cast<const float *>(d.in) + i), d.gain, reduce, reducing, a, y);
                if (MODE != 3) d.out[i] = o;
                if (MODE == 0 && d.before) d.before[i] = y;
            }
        } else {
            for (long long i = lo + threadIdx.x; i < hi; i += kGainThreads) {
                const float x = d.pcm_i16 ? static_cast<float>(__ldg(static_cast<const short *>(d.in) + i)) * 3.0517578125e-05f
                                          : __ldg(static_cast<const float *>(d.in) + i);
                float y;
                const float o = gain_one<MODE>(x, d.gain, reduce, reducing, a, y);
                if (MODE != 3) d.out[i] = o;
                if (MODE == 0 && d.before) d.before[i] = y;
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (MODE == 3) {
        const float m = warp_max(a.mx_before);
        if (lane == 0 && m > 0.0f) atomicMax(group_peak_out + d.group, __float_as_uint(m));
        return;
    }
    const double s = warp_sum_f64(a.ss);
    const float mo = warp_max(a.mx_out);
    if (lane == 0) {
        red_s[warp] = s;
        if (mo > 0.0f) atomicMax(&outs[blockIdx.y].abs_max_bits, __float_as_uint(mo));
    }
    if (MODE == 0) {
        const float mb = warp_max(a.mx_before);
        const unsigned c = __reduce_add_sync(0xffffffffu, a.cnt);
        if (lane == 0) {
            if (mb > 0.0f) atomicMax(&outs[blockIdx.y].before_max_bits, __float_as_uint(mb));
            if (c) atomicAdd(&outs[blockIdx.y].reduction_cnt, static_cast<unsigned long long>(c));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kGainThreads / 32; w++) t += red_s[w];  // fixed order
        part_ss[static_cast<size_t>(blockIdx.y) * chunks_per_ch + blockIdx.x] = t;
    }
}

// one warp per channel: partial sums in ascending chunk order (lane-strided, then a fixed shuffle tree)
__global__ void gain_final_kernel(const double *__restrict__ part_ss, int chunks_per_ch, const GainDesc *__restrict__ descs,
                                  GainOut *__restrict__ outs) {
    const int ch = blockIdx.x, lane = threadIdx.x;
    const long long n_chunks = (descs[ch].len + kGainChunk - 1) / kGainChunk;
    double s = 0.0;
    for (long long c = lane; c < n_chunks; c += 32) s += part_ss[static_cast<size_t>(ch) * chunks_per_ch + c];
    s = warp_sum_f64(s);
    if (lane == 0) outs[ch].sum_squares = static_cast<float>(s);
}

template <int MODE>
cudaError_t launch_mode(const GainDesc *d_descs, int n, int chunks, const unsigned *d_group_peak, double *d_part_ss,
                        GainOut *d_outs, unsigned *d_group_peak_out, cudaStream_t st) {
    for (int c0 = 0; c0 < n; c0 += 65535) {
        const int nc = n - c0 < 65535 ? n - c0 : 65535;
        dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(nc));
        gain_kernel<MODE><<<grid, kGainThreads, 0, st>>>(d_descs + c0, chunks, d_group_peak,
                                                        d_part_ss ? d_part_ss + static_cast<size_t>(c0) * chunks : nullptr,
                                                        d_outs ? d_outs + c0 : nullptr, d_group_peak_out);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (MODE != 3) {
            gain_final_kernel<<<nc, 32, 0, st>>>(d_part_ss + static_cast<size_t>(c0) * chunks, chunks, d_descs + c0, d_outs + c0);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

}  // namespace

long long gain_chunks(long long max_len) { return max_len > 0 ? (max_len + kGainChunk - 1) / kGainChunk : 1; }

cudaError_t launch_gain_peak(const GainDesc *d_descs, int n, long long max_len, unsigned *d_group_peak, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    return launch_mode<3>(d_descs, n, static_cast<int>(gain_chunks(max_len)), nullptr, nullptr, nullptr, d_group_peak, st);
}

cudaError_t launch_gain_apply(const GainDesc *d_descs, int n, long long max_len, int mode, const unsigned *d_group_peak,
                              double *d_part_ss, GainOut *d_outs, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int chunks = static_cast<int>(gain_chunks(max_len));
    switch (mode) {
    case 0: return launch_mode<0>(d_descs, n, chunks, nullptr, d_part_ss, d_outs, nullptr, st);
    case 1: return launch_mode<1>(d_descs, n, chunks, d_group_peak, d_part_ss, d_outs, nullptr, st);
    default: return launch_mode<2>(d_descs, n, chunks, nullptr, d_part_ss, d_outs, nullptr, st);
    }
}

}  // namespace thb
