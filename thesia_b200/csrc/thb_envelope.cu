// thb_envelope.cu -- K5: waveform min/max/mean envelope tiles, written by the device directly in
// the reference's wire format (encode_waveform_tile, render_tiles.rs:232-279):
//   tile t of level L (samples_per_bin = 2^L) covers samples [t*1024*spb, min(N, (t+1)*1024*spb));
//   24-byte header {u64 revision, u32 bin_count, u32 spb, u32 tile_index, u32 0}, then per bin
//   {f32 min, f32 max, f32 mean}.  A full tile is 24 + 12*1024 = 12312 bytes; tiles of one level are
//   stored back to back.
// Pure HBM streaming: 4 B read per sample, 12 B written per bin.  One warp owns a run of 128
// consecutive samples per step (32 lanes x float4, a fully coalesced 512 B request); bins smaller
// than 128 samples are finished with a segmented xor-shuffle, larger ones accumulate across steps.
// min / max are exact; mean = sum / len in f32 (the reference's own sum is SIMD-lane ordered and
// alignment dependent, simd.rs:594-619, so only min/max are bit-pinned -- see DESIGN.md).
#include "thb_device.cuh"
#include "thb_kernels.cuh"

namespace thb {
namespace {

constexpr unsigned kTileBins = 1024;
constexpr unsigned kTileBytes = 24 + 12 * kTileBins;
constexpr int kWarpsPerBlock = 8;

struct Acc {
    float mn, mx, sum;
};
__device__ __forceinline__ Acc acc_init() { return {CUDART_INF_F, -CUDART_INF_F, 0.0f}; }
__device__ __forceinline__ void acc_add(Acc &a, float v) {
    a.mn = fminf(a.mn, v);
    a.mx = fmaxf(a.mx, v);
    a.sum += v;
}
__device__ __forceinline__ void acc_merge_xor(Acc &a, int offset) {
    a.mn = fminf(a.mn, __shfl_xor_sync(0xffffffffu, a.mn, offset));
    a.mx = fmaxf(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, offset));
    a.sum += __shfl_xor_sync(0xffffffffu, a.sum, offset);
}

__device__ __forceinline__ void store_bin(uint8_t *out, unsigned long long bin, const Acc &a, float count) {
    const unsigned long long tile = bin / kTileBins;
    const unsigned b = static_cast<unsigned>(bin % kTileBins);
    float *o = reinterpret_cast<float *>(out + tile * kTileBytes + 24 + 12ull * b);
    o[0] = a.mn;
    o[1] = a.mx;
    o[2] = a.sum / count;
}

// level_log2 = log2(samples per bin), 0..40.  Each warp processes `bins_per_warp` bins (large bins)
// or one 128-sample run after another (small bins).
__global__ void __launch_bounds__(kWarpsPerBlock * 32) envelope_kernel(const EnvDesc *__restrict__ descs,
                                                                       unsigned level, unsigned long long revision,
                                                                       unsigned tile_begin,
                                                                       unsigned long long s_begin,
                                                                       unsigned long long s_limit) {
    const EnvDesc d = descs[blockIdx.y];
    const unsigned long long spb = 1ull << level;
    const unsigned long long n = static_cast<unsigned long long>(d.len);
    // the requested window of the channel: [s_begin, min(len, s_limit))
    const unsigned long long s_end = n < s_limit ? n : s_limit;
    const unsigned long long total_bins = s_end > s_begin ? (s_end - s_begin + spb - 1) / spb : 0;
    if (total_bins == 0) return;
    const unsigned long long n_tiles = (total_bins + kTileBins - 1) / kTileBins;
    const int lane = threadIdx.x & 31;
    const unsigned long long gwarp = static_cast<unsigned long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    const unsigned long long n_warps = static_cast<unsigned long long>(gridDim.x) * kWarpsPerBlock;
    uint8_t *out = d.out;

    // headers: one thread per tile
    for (unsigned long long t = gwarp * 32 + lane; t < n_tiles; t += n_warps * 32) {
        const unsigned long long bins_here = min(static_cast<unsigned long long>(kTileBins), total_bins - t * kTileBins);
        uint32_t *h = reinterpret_cast<uint32_t *>(out + t * kTileBytes);
        h[0] = static_cast<uint32_t>(revision);
        h[1] = static_cast<uint32_t>(revision >> 32);
        h[2] = static_cast<uint32_t>(bins_here);
        h[3] = static_cast<uint32_t>(spb < 0xffffffffull ? spb : 0xffffffffull);
        h[4] = tile_begin + static_cast<uint32_t>(t);
        h[5] = 0u;
    }
    const float *pcm = d.pcm;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;

    if (level >= 7) {
        // one warp per bin, 128 samples per step
        for (unsigned long long bin = gwarp; bin < total_bins; bin += n_warps) {
            const unsigned long long b0 = s_begin + bin * spb;
            const unsigned long long b1 = min(b0 + spb, s_end);
            Acc a = acc_init();
            unsigned long long s = b0 + 4ull * lane;
            if (vec_ok) {
                // b0 is a multiple of 128 samples -> 16-byte aligned float4 loads; four of them in flight per lane
                // (2 KB per warp and step) so that a 64-warp SM keeps enough bytes outstanding to cover HBM latency
                for (; s + 3 * 128 + 3 < b1; s += 512) {
                    const float4 v0 = __ldg(reinterpret_cast<const float4 *>(pcm + s));
                    const float4 v1 = __ldg(reinterpret_cast<const float4 *>(pcm + s + 128));
                    const float4 v2 = __ldg(reinterpret_cast<const float4 *>(pcm + s + 256));
                    const float4 v3 = __ldg(reinterpret_cast<const float4 *>(pcm + s + 384));
                    acc_add(a, v0.x); acc_add(a, v0.y); acc_add(a, v0.z); acc_add(a, v0.w);
                    acc_add(a, v1.x); acc_add(a, v1.y); acc_add(a, v1.z); acc_add(a, v1.w);
                    acc_add(a, v2.x); acc_add(a, v2.y); acc_add(a, v2.z); acc_add(a, v2.w);
                    acc_add(a, v3.x); acc_add(a, v3.y); acc_add(a, v3.z); acc_add(a, v3.w);
                }
                for (; s + 3 < b1; s += 128) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(pcm + s));
                    acc_add(a, v.x); acc_add(a, v.y); acc_add(a, v.z); acc_add(a, v.w);
                }
                for (unsigned long long q = s; q < b1 && q < s + 4; q++) acc_add(a, __ldg(pcm + q));
            } else {
                for (; s < b1; s += 128)
                    for (unsigned long long q = s; q < b1 && q < s + 4; q++) acc_add(a, __ldg(pcm + q));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc_merge_xor(a, o);
            if (lane == 0) store_bin(out, bin, a, static_cast<float>(b1 - b0));
        }
    } else {
        // small bins: a warp step covers 128 samples = 128/spb bins
        const unsigned long long n_steps = (s_end - s_begin + 127) / 128;
        for (unsigned long long step = gwarp; step < n_steps; step += n_warps) {
            const unsigned long long s = s_begin + step * 128 + 4ull * lane;
            float v[4];
            if (vec_ok && s + 3 < s_end) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(pcm + s));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) v[e] = (s + e < s_end) ? __ldg(pcm + s + e) : 0.0f;
            }
            const unsigned long long rel = s - s_begin;  // sample offset inside the window
            if (level == 0) {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (s + e < s_end) {
                        Acc a = {v[e], v[e], v[e]};
                        store_bin(out, rel + e, a, 1.0f);
                    }
            } else if (level == 1) {
#pragma unroll
                for (int e = 0; e < 4; e += 2)
                    if (s + e < s_end) {
                        Acc a = acc_init();
                        acc_add(a, v[e]);
                        const bool two = s + e + 1 < s_end;
                        if (two) acc_add(a, v[e + 1]);
                        store_bin(out, (rel + e) >> 1, a, two ? 2.0f : 1.0f);
                    }
            } else {
                Acc a = acc_init();
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (s + e < s_end) acc_add(a, v[e]);
                const int group = 1 << (level - 2);  // lanes per bin: 1, 2, 4, 8, 16
                for (int o = 1; o < group; o <<= 1) acc_merge_xor(a, o);
                if ((lane & (group - 1)) == 0 && s < s_end) {
                    const unsigned long long bin = rel >> level;
                    const unsigned long long b0 = s_begin + bin * spb;
                    const unsigned long long b1 = min(b0 + spb, s_end);
                    store_bin(out, bin, a, static_cast<float>(b1 - b0));
                }
            }
        }
    }
}

// ---- deterministic synthetic PCM (integer arithmetic only; thesia_b200/synth.py is the numpy twin) ----
__device__ __forceinline__ int para_sine(unsigned phase) {
    // parabolic "sine": phase is a 32-bit turn fraction; result in [-65536, 65536]
    const unsigned x = (phase >> 15) & 0xffffu;
    const int y = static_cast<int>((static_cast<unsigned long long>(x) * (65536ull - x)) >> 14);
    return (phase >> 31) ? -y : y;
}
__device__ __forceinline__ unsigned mix32(unsigned h) {
    h ^= h >> 15; h *= 0x85ebca77u;
    h ^= h >> 13; h *= 0xc2b2ae3du;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ int synth_base(unsigned long long n, unsigned long long len, unsigned sr,
                                          unsigned track, unsigned flags) {
    // tone: f0 = 55 Hz + 13.75 Hz * (track % 61)
    const unsigned long long f0_mhz = 55000ull + 13750ull * (track % 61u);
    const unsigned long long inc0 = (f0_mhz << 32) / (1000ull * sr);
    const unsigned ph0 = static_cast<unsigned>(n * inc0);
    // linear chirp 50 Hz -> 0.45 sr over the file
    const unsigned long long inc_a = (50ull << 32) / sr;
    const unsigned long long inc_b = 1932735283ull;  // floor(0.45 * 2^32)
    const unsigned long long dinc = inc_b - inc_a;
    const unsigned long long n2 = n * n;             // n < 2^32
    const unsigned long long two_len = 2ull * len;
    const unsigned long long t_int = n2 / two_len, t_rem = n2 % two_len;
    const unsigned ph1 = static_cast<unsigned>(inc_a * n + dinc * t_int + (dinc * t_rem) / two_len);
    const unsigned h = mix32(static_cast<unsigned>(n) * 0x9e3779b1u + track * 0x7f4a7c15u + 0x7e51au);
    const int a0 = (8192 * para_sine(ph0)) >> 16;
    const int a1 = (3277 * para_sine(ph1)) >> 16;
    const int nz = ((static_cast<int>(h >> 16) - 32768) * 1638) >> 15;
    int v = a0 + a1 + nz;
    if ((flags & 2u) && n >= sr && n < 2ull * sr) v = 0;
    if (flags & 1u) v *= 32;
    return v;
}
__global__ void synth_pcm_kernel(float *out, unsigned long long len, unsigned sr, unsigned track,
                                 unsigned channel, unsigned flags) {
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long n = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; n < len;
         n += stride) {
        int v;
        if (channel == 0) v = synth_base(n, len, sr, track, flags);
        else v = n >= 7 ? static_cast<int>((static_cast<long long>(synth_base(n - 7, len, sr, track, flags)) * 26214) >> 15) : 0;
        out[n] = static_cast<float>(v) / 32768.0f;
    }
}

}  // namespace

cudaError_t launch_envelope(const EnvDesc *d_descs, int n, long long max_len, uint32_t level,
                            uint64_t revision, uint32_t tile_begin, uint32_t tile_count, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (level > 40) level = 40;
    const unsigned long long spb = 1ull << level;
    const unsigned long long tile_samples = spb * kTileBins;  // <= 2^50
    // window [s_begin, s_limit) in samples; tile_count == 0 means "to the end of the channel"
    const unsigned long long s_begin = static_cast<unsigned long long>(tile_begin) * tile_samples;  // < 2^82?  no: 2^32*2^50
    unsigned long long s_limit = ~0ull;
    if (tile_count != 0) {
        const unsigned __int128 lim = static_cast<unsigned __int128>(tile_begin + static_cast<unsigned long long>(tile_count)) * tile_samples;
        s_limit = lim > static_cast<unsigned __int128>(~0ull) ? ~0ull : static_cast<unsigned long long>(lim);
    }
    if (static_cast<unsigned __int128>(tile_begin) * tile_samples > static_cast<unsigned __int128>(~0ull)) return cudaSuccess;
    unsigned long long span = static_cast<unsigned long long>(max_len);
    if (s_limit < span) span = s_limit;
    span = span > s_begin ? span - s_begin : 0;
    if (span == 0) return cudaSuccess;
    // work items: bins (level >= 7) or 128-sample runs (level < 7); one warp each, several per warp
    unsigned long long items = level >= 7 ? (span + spb - 1) / spb : (span + 127) / 128;
    unsigned long long blocks = (items + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const unsigned long long cap = 148ull * 8 * 4;  // a few waves of 8 resident CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    for (int t0 = 0; t0 < n; t0 += 65535) {
        const int nt = n - t0 < 65535 ? n - t0 : 65535;
        dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(nt));
        envelope_kernel<<<grid, kWarpsPerBlock * 32, 0, st>>>(d_descs + t0, level, revision, tile_begin, s_begin,
                                                              s_limit);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_synth_pcm(float *d_out, unsigned long long len, uint32_t sr, uint32_t track,
                             uint32_t channel, uint32_t flags, cudaStream_t st) {
    if (len == 0) return cudaSuccess;
    unsigned long long blocks = (len + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    synth_pcm_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(d_out, len, sr, track, channel, flags);
    return cudaGetLastError();
}

}  // namespace thb
