// thb_stft_fast.cu -- K1/K2/K3 for n_fft == 2048 (the 40 ms / 42.7 ms windows at 44.1 and 48 kHz):
// one WARP transforms one frame, entirely in registers, with a single shared-memory transpose.
//
//   frame  : 2048 real samples packed as 1024 complex z[m] = x[2m] + i x[2m+1]
//            (perform_stft / to_windowed_frames, stft.rs:16-149; centred zero pad stft.rs:35-39)
//   FFT    : 1024 = 32 x 32.  Lane n2 holds z[32 n1 + n2], n1 = 0..31 (coalesced 8-byte loads of the
//            PCM, multiplied by the zero-padded window from shared memory), runs a 32-point DFT in
//            registers (radix 4 x 8, compile-time twiddles), multiplies by W_1024^(n2 k1) and writes
//            row k1 of a 32 x 33 float2 tile; lane k1 reads its row back and runs the second 32-point
//            DFT, leaving Z[k1 + 32 k2].  One conflict-free transpose, no block barrier.
//   split  : X[k], X[1024-k] of the real-input transform (stft.rs:44-48) from Z[k], Z[1024-k]; the
//            partner value lives in lane (32 - k1) & 31 and comes over one shuffle pair; every lane
//            resolves 16 (k, 1024-k) pairs = 32 bins.
//   |X|    : Complex::norm (spectrogram.rs:200); linear scale: dB straight from |X|^2.
//   mel    : band-major sparse triangular filters on |X| staged in the warp's tile
//            (= linspec.dot(mel_fb), spectrogram.rs:207, without the zero weights)
//   dB     : log10 * 20 (decibel.rs:198-202) through MUFU.LG2; running {max, -min} per channel
//            (find_min_max, mod.rs:169-178), one atomic pair per CTA.
// The 0.5 of the real-split is folded into the window table (exact power-of-two scaling).
#include "thb_stft2048.cuh"

namespace thb {

namespace {

using namespace k2048;

constexpr int kWarps = 8;        // frames in flight per CTA
constexpr int kTileFrames = 64;  // consecutive frames of one channel per CTA (grid mode)

// a * W_32^J,  W_32 = exp(-2 pi i / 32)
template <int J>
__device__ __forceinline__ float2 mul_w32(float2 a) {
    constexpr int j = J & 31;
    if constexpr (j == 0) {
        return a;
    } else if constexpr (j == 8) {
        return make_float2(a.y, -a.x);
    } else if constexpr (j == 16) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (j == 24) {
        return make_float2(-a.y, a.x);
    } else {
        constexpr float c = kC32[j], s = kS32[j];
        return make_float2(fmaf(a.y, s, a.x * c), fmaf(-a.x, s, a.y * c));
    }
}

__device__ __forceinline__ void dft8r(float2 &v0, float2 &v1, float2 &v2, float2 &v3, float2 &v4, float2 &v5,
                                      float2 &v6, float2 &v7) {
    const float h = 0.70710678118654752440f;
    float2 t0 = cadd(v0, v4), u0 = csub(v0, v4);
    float2 t1 = cadd(v1, v5), u1 = csub(v1, v5);
    float2 t2 = cadd(v2, v6), u2 = csub(v2, v6);
    float2 t3 = cadd(v3, v7), u3 = csub(v3, v7);
    u1 = make_float2(h * (u1.x + u1.y), h * (u1.y - u1.x));
    u2 = cmul_mi(u2);
    u3 = make_float2(h * (u3.y - u3.x), -h * (u3.x + u3.y));
    dft4(t0, t1, t2, t3);
    dft4(u0, u1, u2, u3);
    v0 = t0; v2 = t1; v4 = t2; v6 = t3;
    v1 = u0; v3 = u1; v5 = u2; v7 = u3;
}

// In-register 32-point forward DFT.  Output X[k] is left in v[perm32(k)].
template <int B>
__device__ __forceinline__ void dft32_col(float2 (&v)[32]) {
    dft4(v[B], v[B + 8], v[B + 16], v[B + 24]);
    v[B + 8] = mul_w32<B>(v[B + 8]);
    v[B + 16] = mul_w32<2 * B>(v[B + 16]);
    v[B + 24] = mul_w32<3 * B>(v[B + 24]);
}

__device__ __forceinline__ void dft32(float2 (&v)[32]) {
    // n = b + 8a, k = q + 4r:  W32^(nk) = W4^(aq) W32^(bq) W8^(br)
    dft32_col<0>(v); dft32_col<1>(v); dft32_col<2>(v); dft32_col<3>(v);
    dft32_col<4>(v); dft32_col<5>(v); dft32_col<6>(v); dft32_col<7>(v);
#pragma unroll
    for (int q = 0; q < 4; q++)
        dft8r(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6],
              v[8 * q + 7]);
}

struct FastSmem {
    float *wpad;         // [2048]   0.5 * window, zero outside the taps
    float2 *tw1;         // [31][32] W_1024^(lane * k1), k1 = 1..31
    float2 *tw2;         // [16][32] W_2048^(k_own(j, lane))
    float *tiles;        // [kWarps][tile_floats]
    const uint32_t *ms;  // MelItems blob
};

// floats per warp tile: the float2 transpose tile, or magnitudes + mel partial sums, whichever is larger
__host__ __device__ inline int fast_tile_floats(const PlanDev &p) {
    const int t = tile_elems(p);
    return t > 2 * 32 * kRow ? t : 2 * 32 * kRow;
}

__device__ __forceinline__ FastSmem carve(unsigned char *raw, int tile_floats) {
    FastSmem s;
    s.wpad = reinterpret_cast<float *>(raw);
    s.tw1 = reinterpret_cast<float2 *>(s.wpad + 2048);
    s.tw2 = s.tw1 + 31 * 32;
    s.tiles = reinterpret_cast<float *>(s.tw2 + 16 * 32);
    s.ms = reinterpret_cast<const uint32_t *>(s.tiles + kWarps * tile_floats);
    return s;
}

// LIST = false: grid (tiles, descriptors).  LIST = true: a persistent grid walks the rescue list the
// frame-pair kernel left behind (thb_stft_pair.cu) -- normally empty, so every CTA exits at once.
template <bool MEL, bool LIST>
__global__ void __launch_bounds__(kWarps * 32, 2) stft2048_kernel(const PlanDev p,
                                                                   const TrackDesc *__restrict__ tracks,
                                                                   const RescueList rescue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red_max[kWarps], red_nmin[kWarps];

    const int tile_frames = LIST ? static_cast<int>(rescue.tile_frames) : kTileFrames;
    unsigned n_items = 1;
    if (LIST) {
        pdl_wait();   // launched under the tail of the frame-pair kernel that fills the list
        n_items = min(*rescue.count, rescue.capacity);
        if (blockIdx.x >= n_items) return;
    } else if (static_cast<long long>(blockIdx.x) * kTileFrames >= tracks[blockIdx.y].n_frames) {
        return;
    }

    const int tile_floats = fast_tile_floats(p);
    const FastSmem sm = carve(smem_raw, tile_floats);
    // ---- tables -> shared memory; tiles start out as zeros (the mel walk reads padding x 0) ----
    for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(sm.wpad)[i] = __ldg(reinterpret_cast<const float4 *>(p.fast_wpad) + i);
    for (int i = threadIdx.x; i < (31 * 32 + 16 * 32) / 2; i += blockDim.x)
        reinterpret_cast<float4 *>(sm.tw1)[i] = __ldg(reinterpret_cast<const float4 *>(p.fast_tw) + i);
    for (int i = threadIdx.x; i < kWarps * tile_floats / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(sm.tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MEL) {
        for (int i = threadIdx.x; i < p.mi_words / 4; i += blockDim.x)
            reinterpret_cast<uint4 *>(const_cast<uint32_t *>(sm.ms))[i] = __ldg(reinterpret_cast<const uint4 *>(p.mi_blob) + i);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *tile = reinterpret_cast<float2 *>(sm.tiles + warp * tile_floats);
    float *mag = reinterpret_cast<float *>(tile) + kMagBase;
    float *part = reinterpret_cast<float *>(tile) + part_base(p);
    const int half = p.win / 2;
    const int partner = (32 - lane) & 31;
    const int lane32 = lane ? lane : 32;

    for (unsigned item = LIST ? blockIdx.x : 0; item < n_items; item += LIST ? gridDim.x : 1) {
        const uint2 it = LIST ? rescue.items[item] : make_uint2(blockIdx.y, blockIdx.x);
        const TrackDesc d = tracks[it.x];
        const long long f_begin = static_cast<long long>(it.y) * tile_frames;
        const long long f_end = min(f_begin + tile_frames, d.n_frames);
        float lmax = -CUDART_INF_F, lnmin = -CUDART_INF_F;

        for (long long f = f_begin + warp; f < f_end; f += kWarps) {
            float2 v[32];
            // ---- load + window: v[n1] = z[32 n1 + lane] ----
            const long long tap0 = (d.frame_begin + f) * p.hop - half;  // file index of window tap 0
            const long long first = tap0 - p.pad_left;                  // file index of FFT position 0
            const bool interior = tap0 >= 0 && tap0 + p.win <= d.full_len && first >= d.pcm_offset &&
                                  first + 2048 <= d.pcm_offset + d.slice_len;
            if (interior) {
                const float *src = d.pcm + (first - d.pcm_offset) + 2 * lane;
                if (d.pcm_i16) {
                    const long long at = (first - d.pcm_offset) + 2 * lane;
#pragma unroll
                    for (int n1 = 0; n1 < 32; n1++)
                        v[n1] = make_float2(pcm_sample(d, at + 64 * n1), pcm_sample(d, at + 64 * n1 + 1));
                } else if ((reinterpret_cast<uintptr_t>(src) & 7) == 0) {
#pragma unroll
                    for (int n1 = 0; n1 < 32; n1++) v[n1] = __ldg(reinterpret_cast<const float2 *>(src + 64 * n1));
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < 32; n1++) v[n1] = make_float2(__ldg(src + 64 * n1), __ldg(src + 64 * n1 + 1));
                }
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                    v[n1].x *= w.x;
                    v[n1].y *= w.y;
                }
            } else {
                // file edges: numpy-style reflect (utils.rs:111-137), taps outside the window are zero;
                // staged through the warp's tile so that v[] keeps compile-time indices
                float *stage = reinterpret_cast<float *>(tile);
                stage_edge_frame<2048>(d, p.pad_left, p.win, tap0, sm.wpad, stage, lane);
                __syncwarp();
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) v[n1] = *reinterpret_cast<const float2 *>(stage + 64 * n1 + 2 * lane);
                __syncwarp();
            }
            // ---- pass 1: DFT over n1, twiddle, transpose ----
            dft32(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; k1++) {
                float2 t = v[perm32(k1)];
                if (k1) {  // same operation order as the packed kernel's cmul_s: the two kernels agree bit for bit
                    const float2 w = sm.tw1[(k1 - 1) * 32 + lane];
                    t = make_float2(fmaf(t.y, -w.y, t.x * w.x), fmaf(t.x, w.y, t.y * w.x));
                }
                tile[k1 * kRow + lane] = t;
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) v[n2] = tile[lane * kRow + n2];
            __syncwarp();
            // ---- pass 2: DFT over n2 -> Z[lane + 32 k2] in v[perm32(k2)] ----
            dft32(v);
            // ---- real split: 16 (k, 1024 - k) pairs per lane; |X| or dB ----
            float *orow = d.out + f * p.n_bins;
            float db_off = 0.0f;
            for (int attempt = 0;; attempt++) {
                float smax = 0.0f, fmx = -CUDART_INF_F, fnm = -CUDART_INF_F;  // this attempt's own max / -min
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const float2 own_a = v[perm32(31 - j)], own_b = v[perm32((32 - j) & 31)];
                    const float2 zk = lane ? own_a : own_b;
                    const float2 sup = v[perm32(j)];
                    float2 zn;
                    zn.x = __shfl_sync(0xffffffffu, sup.x, partner);
                    zn.y = __shfl_sync(0xffffffffu, sup.y, partner);
                    const float er = zk.x + zn.x, ei = zk.y - zn.y, dr = zk.x - zn.x, di = zk.y + zn.y;
                    const float2 w = sm.tw2[j * 32 + lane];
                    const float wr = fmaf(di, -w.y, dr * w.x), wi = fmaf(dr, w.y, di * w.x);
                    const float ar = er + wi, ai = ei - wr, br = er - wi, bi = ei + wr;
                    const float sa = fmaf(ar, ar, ai * ai), sb = fmaf(br, br, bi * bi);
                    smax = fmaxf(smax, fmaxf(sa, sb));
                    const int k_own = lane32 + 32 * (31 - j), k_par = 1024 - k_own;
                    if (MEL) {
                        mag[k_own] = sqrt_ftz(sa);
                        mag[k_par] = sqrt_ftz(sb);
                    } else {
                        const float a = fmaf(kDbPerLog2Pow, lg2_ftz(sa), db_off), b = fmaf(kDbPerLog2Pow, lg2_ftz(sb), db_off);
                        orow[k_own] = a;
                        orow[k_par] = b;
                        fmx = fmaxf(fmx, fmaxf(a, b));
                        fnm = fmaxf(fnm, fmaxf(-a, -b));
                    }
                }
                if (lane == 0) {  // k = 512 pairs with itself: X[512] = conj(2 Z'[512])
                    const float2 z = v[perm32(16)];
                    const float s5 = 4.0f * fmaf(z.x, z.x, z.y * z.y);
                    smax = fmaxf(smax, s5);
                    if (MEL) {
                        mag[512] = sqrt_ftz(s5);
                    } else {
                        const float a = fmaf(kDbPerLog2Pow, lg2_ftz(s5), db_off);
                        orow[512] = a;
                        fmx = fmaxf(fmx, a);
                        fnm = fmaxf(fnm, -a);
                    }
                }
                smax = warp_max(smax);
                const bool tiny = smax < kPowTiny && smax > 0.0f, huge = smax > kPowHuge;
                if (attempt || !(tiny || huge)) {
                    lmax = fmaxf(lmax, fmx);
                    lnmin = fmaxf(lnmin, fnm);
                    break;
                }
                const float sc = tiny ? kRescueUp : kRescueDown;
                db_off = tiny ? -60.0f * kDbPerLog2Amp : 60.0f * kDbPerLog2Amp;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    v[i].x *= sc;
                    v[i].y *= sc;
                }
            }
            if (MEL) {
                __syncwarp();
                const MelView mv(sm.ms);
                if (!mv.direct) {
                    mel_walk4<float>(mv, mag, part, lane);
                    __syncwarp();
                }
                for (int r = 0; 32 * r < mv.n_mel; r++) {
                    const int m = 32 * r + lane;
                    const float acc = mv.direct ? mel_direct<float>(mv, mag, r, lane) : mel_band4<float>(mv, part, r, lane);
                    if (m >= mv.n_mel) continue;
                    // exact zero stays -inf; db_off only shifts finite values
                    const float db = fmaf(kDbPerLog2Amp, lg2_ftz(acc), db_off);
                    orow[m] = db;
                    lmax = fmaxf(lmax, db);
                    lnmin = fmaxf(lnmin, -db);
                }
                __syncwarp();
            }
        }
        // ---- per-channel {max, -min}: warp shuffle -> shared -> one atomic pair per CTA ----
        lmax = warp_max(lmax);
        lnmin = warp_max(lnmin);
        if (lane == 0) {
            red_max[warp] = lmax;
            red_nmin[warp] = lnmin;
        }
        __syncthreads();
        if (warp == 0) {
            float a = lane < kWarps ? red_max[lane] : -CUDART_INF_F;
            float b = lane < kWarps ? red_nmin[lane] : -CUDART_INF_F;
            a = warp_max(a);
            b = warp_max(b);
            if (lane == 0) {
                atomic_max_float(&d.minmax[0], a);
                atomic_max_float(&d.minmax[1], b);
            }
        }
        __syncthreads();  // red_max / red_nmin are reused by the next item
    }
}

size_t fast_smem_bytes(const PlanDev &p) {
    return sizeof(float) * 2048 + sizeof(float2) * (31 * 32 + 16 * 32) + sizeof(float) * kWarps * fast_tile_floats(p) +
           sizeof(uint32_t) * static_cast<size_t>(p.n_mel ? p.mi_words : 0);
}

}  // namespace

bool stft_fast_supported(const PlanDev &p) {
    if (p.n_fft != 2048 || !p.fast_wpad || !p.fast_tw) return false;
    if (p.n_mel && (!p.mi_blob || p.mi_min_start < -(kMagBase - 1))) return false;
    return fast_smem_bytes(p) <= 112 * 1024;
}

cudaError_t launch_stft_fast(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, long long max_frames,
                             int /*sm_count*/, cudaStream_t st) {
    if (n_tracks <= 0 || max_frames <= 0) return cudaSuccess;
    const size_t smem = fast_smem_bytes(plan);
    auto kern = plan.n_mel ? stft2048_kernel<true, false> : stft2048_kernel<false, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const long long tiles = (max_frames + kTileFrames - 1) / kTileFrames;
    for (int t0 = 0; t0 < n_tracks; t0 += 65535) {
        const int nt = min(65535, n_tracks - t0);
        dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(nt));
        kern<<<grid, kWarps * 32, smem, st>>>(plan, d_tracks + t0, RescueList{});
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_stft_fast_list(const PlanDev &plan, const TrackDesc *d_tracks, RescueList rescue, int sm_count,
                                  cudaStream_t st) {
    const size_t smem = fast_smem_bytes(plan);
    auto kern = plan.n_mel ? stft2048_kernel<true, true> : stft2048_kernel<false, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return launch_pdl(kern, dim3(2 * sm_count), dim3(kWarps * 32), smem, st, plan, d_tracks, rescue);
}

}  // namespace thb
