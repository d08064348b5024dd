// thb_stats.cu -- per-channel level statistics: sum of squares and absolute maximum of a
// channel's samples, the two reductions StatCalculator::calc runs over the PCM (dynamics/stats.rs:56-85 through
// sum_squares, simd.rs:113-134,820-832, and abs_max, simd.rs:161-183,935-937).
//
// HBM-bound: 4 bytes (2 for 16-bit PCM) read per sample, nothing written.  The reference compensates its f32 sum
// (Kahan); here every thread accumulates in f64, which is more exact than that, and the partial sums of the CTAs
// are added in a fixed order by a second one-warp-per-channel kernel, so the result does not depend on scheduling.
#include "thb_device.cuh"
#include "thb_kernels.cuh"

namespace thb {
namespace {

constexpr int kStatThreads = 256;
constexpr long long kStatChunk = 1 << 18;  // samples per CTA: 1 MB of f32

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kStatThreads) stats_partial_kernel(const TrackDesc *__restrict__ descs, int chunks_per_ch,
                                                                     double *__restrict__ part_ss, float *__restrict__ part_mx) {
    __shared__ double red_s[kStatThreads / 32];
    __shared__ float red_m[kStatThreads / 32];
    const TrackDesc d = descs[blockIdx.y];
    const long long lo = static_cast<long long>(blockIdx.x) * kStatChunk;
    const long long hi = min(lo + kStatChunk, d.slice_len);
    double ss = 0.0;
    float mx = 0.0f;
    if (lo < hi) {
        if (!d.pcm_i16 && (reinterpret_cast<uintptr_t>(d.pcm) & 15) == 0) {
            // chunk starts are multiples of 2^18 samples: 16-byte aligned whenever the channel is
            const float4 *p = reinterpret_cast<const float4 *>(d.pcm + lo);
            const long long n4 = (hi - lo) >> 2;
            for (long long i = threadIdx.x; i < n4; i += kStatThreads) {
                const float4 v = __ldg(p + i);
                ss += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y;
                ss += static_cast<double>(v.z) * v.z + static_cast<double>(v.w) * v.w;
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
            }
            for (long long i = lo + 4 * n4 + threadIdx.x; i < hi; i += kStatThreads) {
                const float v = __ldg(d.pcm + i);
                ss += static_cast<double>(v) * v;
                mx = fmaxf(mx, fabsf(v));
            }
        } else {
            for (long long i = lo + threadIdx.x; i < hi; i += kStatThreads) {
                const float v = pcm_sample(d, i);
                ss += static_cast<double>(v) * v;
                mx = fmaxf(mx, fabsf(v));
            }
        }
    }
    ss = warp_sum_f64(ss);
    mx = warp_max(mx);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red_s[warp] = ss;
        red_m[warp] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        float m = 0.0f;
        for (int w = 0; w < kStatThreads / 32; w++) {  // fixed order
            s += red_s[w];
            m = fmaxf(m, red_m[w]);
        }
        part_ss[static_cast<size_t>(blockIdx.y) * chunks_per_ch + blockIdx.x] = s;
        part_mx[static_cast<size_t>(blockIdx.y) * chunks_per_ch + blockIdx.x] = m;
    }
}

// one warp per channel: partial sums in ascending chunk order (lane-strided, then a fixed shuffle tree)
__global__ void stats_final_kernel(const double *__restrict__ part_ss, const float *__restrict__ part_mx, int chunks_per_ch,
                                   const TrackDesc *__restrict__ descs, float *__restrict__ out_ss, float *__restrict__ out_mx) {
    const int ch = blockIdx.x, lane = threadIdx.x;
    const long long n_chunks = (descs[ch].slice_len + kStatChunk - 1) / kStatChunk;
    double s = 0.0;
    float m = 0.0f;
    for (long long c = lane; c < n_chunks; c += 32) {
        s += part_ss[static_cast<size_t>(ch) * chunks_per_ch + c];
        m = fmaxf(m, part_mx[static_cast<size_t>(ch) * chunks_per_ch + c]);
    }
    s = warp_sum_f64(s);
    m = warp_max(m);
    if (lane == 0) {
        out_ss[ch] = static_cast<float>(s);
        out_mx[ch] = m;
    }
}

}  // namespace

long long stats_chunks(long long max_len) { return max_len > 0 ? (max_len + kStatChunk - 1) / kStatChunk : 1; }

// d_part_ss / d_part_mx: n * stats_chunks(max_len) elements of scratch; d_out_ss / d_out_mx: n results
cudaError_t launch_channel_stats(const TrackDesc *d_descs, int n, long long max_len, double *d_part_ss, float *d_part_mx,
                                 float *d_out_ss, float *d_out_mx, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int chunks = static_cast<int>(stats_chunks(max_len));
    for (int c0 = 0; c0 < n; c0 += 65535) {
        const int nc = n - c0 < 65535 ? n - c0 : 65535;
        dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(nc));
        stats_partial_kernel<<<grid, kStatThreads, 0, st>>>(d_descs + c0, chunks, d_part_ss + static_cast<size_t>(c0) * chunks,
                                                            d_part_mx + static_cast<size_t>(c0) * chunks);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        stats_final_kernel<<<nc, 32, 0, st>>>(d_part_ss + static_cast<size_t>(c0) * chunks, d_part_mx + static_cast<size_t>(c0) * chunks,
                                              chunks, d_descs + c0, d_out_ss + c0, d_out_mx + c0);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace thb
