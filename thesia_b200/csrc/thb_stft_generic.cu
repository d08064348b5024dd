// thb_stft_generic.cu -- K1/K2/K3, shared-memory path for ANY power-of-two n_fft in [4, 32768].
//
// One CTA transforms one frame at a time (a tile of consecutive frames of one channel):
//   load: reflect-indexed PCM * window, centred zero pad, packed as z[n] = x[2n] + i x[2n+1]
//         (perform_stft / to_windowed_frames, stft.rs:16-149)
//   FFT : in-place decimation-in-frequency radix-8/4/2 passes over shared memory, one barrier per
//         pass; the result is left in digit-reversed order and read back through perm()
//   split: X[k] from Z[k], Z[nc-k]  (real-input FFT = the R2C transform of stft.rs:44-48)
//   |X| : Complex::norm (spectrogram.rs:200)
//   mel : band-major sparse triangular filters on |X| (= linspec.dot(mel_fb), spectrogram.rs:207,
//         without the 99.4 % zero weights)
//   dB  : log10 * 20 (decibel.rs:198-202), running {max, -min} per channel (find_min_max,
//         mod.rs:169-178) accumulated with one atomic pair per CTA.
// The fast register path for n_fft == 2048 lives in thb_stft_fast.cu; this kernel is the general
// one (C4's n_fft 16384, small windows, f_overlap > 1).
#include "thb_device.cuh"
#include "thb_kernels.cuh"

namespace thb {

namespace {

constexpr int kFramesPerTile = 8;

__device__ __forceinline__ int zpad(int i) { return i + (i >> 4); }

template <int RL>
__device__ __forceinline__ void dif_pass(float2 *z, int nc, int L, const float2 *__restrict__ tw,
                                         int n_fft) {
    constexpr int R = 1 << RL;
    const int Lr = L >> RL;               // distance between the R inputs of one butterfly
    const int lr_shift = 31 - __clz(Lr);  // Lr is a power of two
    const int nb = nc >> RL;
    const int twstep = n_fft / L;         // exp(-2 pi i q i / L) = tw[q * i * twstep]
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const int blk = j >> lr_shift;
        const int i = j & (Lr - 1);
        const int base = blk * L + i;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = z[zpad(base + r * Lr)];
        if constexpr (RL == 3) {
            dft8(v);
        } else if constexpr (RL == 2) {
            dft4(v[0], v[1], v[2], v[3]);
        } else {
            dft2(v[0], v[1]);
        }
        if (Lr > 1) {
#pragma unroll
            for (int q = 1; q < R; q++) v[q] = cmul(v[q], __ldg(&tw[q * i * twstep]));
        }
#pragma unroll
        for (int q = 0; q < R; q++) z[zpad(base + q * Lr)] = v[q];
    }
}

// position of Z[k] after the in-place DIF passes (digits of k reversed pass by pass)
__device__ __forceinline__ int dif_perm(int k, const PlanDev &p) {
    int pos = 0, rem = p.nc;
#pragma unroll 1
    for (int s = 0; s < p.n_pass; s++) {
        const int rl = p.radix_log2[s];
        const int q = k & ((1 << rl) - 1);
        k >>= rl;
        rem >>= rl;
        pos += q * rem;
    }
    return pos;
}

__global__ void __launch_bounds__(1024) stft_generic_kernel(const PlanDev p,
                                                            const TrackDesc *__restrict__ tracks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *z = reinterpret_cast<float2 *>(smem_raw);
    float *mag = reinterpret_cast<float *>(z + zpad(p.nc) + 2);
    __shared__ float red_max[32], red_nmin[32];

    const TrackDesc d = tracks[blockIdx.y];
    const long long f_begin = static_cast<long long>(blockIdx.x) * kFramesPerTile;
    if (f_begin >= d.n_frames) return;
    const long long f_end = min(f_begin + kFramesPerTile, d.n_frames);

    float lmax = -CUDART_INF_F, lnmin = -CUDART_INF_F;
    const int half = p.win / 2;

    for (long long f = f_begin; f < f_end; f++) {
        // ---- load + window + zero pad ----
        const long long s0 = (d.frame_begin + f) * p.hop - half;  // file index of tap 0
        const bool interior = (s0 >= 0) && (s0 + p.win <= d.full_len);
        for (int n = threadIdx.x; n < p.nc; n += blockDim.x) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int a = 2 * n + e - p.pad_left;  // tap index
                float x = 0.0f;
                if (a >= 0 && a < p.win) {
                    long long s = s0 + a;
                    if (!interior) s = reflect_index(s, d.full_len);
                    x = pcm_sample(d, s - d.pcm_offset) * __ldg(&p.window[a]);
                }
                v[e] = x;
            }
            z[zpad(n)] = make_float2(v[0], v[1]);
        }
        __syncthreads();
        // ---- FFT passes ----
        int L = p.nc;
        for (int s = 0; s < p.n_pass; s++) {
            const int rl = p.radix_log2[s];
            if (rl == 3) dif_pass<3>(z, p.nc, L, p.twiddle, p.n_fft);
            else if (rl == 2) dif_pass<2>(z, p.nc, L, p.twiddle, p.n_fft);
            else dif_pass<1>(z, p.nc, L, p.twiddle, p.n_fft);
            L >>= rl;
            __syncthreads();
        }
        // ---- real split, magnitude, (linear) dB ----
        float *orow = d.out + f * p.n_bins;
        for (int k = threadIdx.x; k <= p.nc; k += blockDim.x) {
            float xr, xi;
            if (k == 0 || k == p.nc) {
                const float2 z0 = z[zpad(0)];
                xr = (k == 0) ? (z0.x + z0.y) : (z0.x - z0.y);
                xi = 0.0f;
            } else {
                const float2 zk = z[zpad(dif_perm(k, p))];
                const float2 zn = z[zpad(dif_perm(p.nc - k, p))];
                const float er = 0.5f * (zk.x + zn.x), ei = 0.5f * (zk.y - zn.y);
                const float dr = 0.5f * (zk.x - zn.x), di = 0.5f * (zk.y + zn.y);
                const float2 w = __ldg(&p.twiddle[k]);
                const float wr = w.x * dr - w.y * di, wi = w.x * di + w.y * dr;
                xr = er + wi;
                xi = ei - wr;
            }
            const float m = cabs_safe(xr, xi);
            if (p.n_mel == 0) {
                const float db = amp_to_db(m);
                orow[k] = db;
                lmax = fmaxf(lmax, db);
                lnmin = fmaxf(lnmin, -db);
            } else {
                mag[k] = m;
            }
        }
        __syncthreads();
        // ---- sparse mel + dB ----
        if (p.n_mel != 0) {
            for (int m = threadIdx.x; m < p.n_mel; m += blockDim.x) {
                const uint32_t k0 = __ldg(&p.mel_k0[m]);
                const uint32_t p0 = __ldg(&p.mel_ptr[m]), p1 = __ldg(&p.mel_ptr[m + 1]);
                float acc = 0.0f;
                for (uint32_t i = p0; i < p1; i++) acc = fmaf(__ldg(&p.mel_w[i]), mag[k0 + (i - p0)], acc);
                const float db = amp_to_db(acc);
                orow[m] = db;
                lmax = fmaxf(lmax, db);
                lnmin = fmaxf(lnmin, -db);
            }
            __syncthreads();
        }
    }
    // ---- per-channel {max, -min}: warp shuffle -> shared -> one atomic pair per CTA ----
    lmax = warp_max(lmax);
    lnmin = warp_max(lnmin);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red_max[warp] = lmax;
        red_nmin[warp] = lnmin;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        float a = lane < nw ? red_max[lane] : -CUDART_INF_F;
        float b = lane < nw ? red_nmin[lane] : -CUDART_INF_F;
        a = warp_max(a);
        b = warp_max(b);
        if (lane == 0) {
            atomic_max_float(&d.minmax[0], a);
            atomic_max_float(&d.minmax[1], b);
        }
    }
}

}  // namespace

cudaError_t launch_stft_generic(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks,
                                long long max_frames, cudaStream_t st) {
    if (n_tracks <= 0 || max_frames <= 0) return cudaSuccess;
    int threads = plan.nc / 8;
    if (threads < 32) threads = 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = sizeof(float2) * (static_cast<size_t>(plan.nc) + (plan.nc >> 4) + 2) +
                        sizeof(float) * (static_cast<size_t>(plan.n_freq) + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(stft_generic_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    const long long tiles = (max_frames + kFramesPerTile - 1) / kFramesPerTile;
    // grid.x carries the tiles (up to 2^31-1), grid.y the channels (<= 65535 per launch)
    for (int t0 = 0; t0 < n_tracks; t0 += 65535) {
        const int nt = min(65535, n_tracks - t0);
        dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(nt));
        stft_generic_kernel<<<grid, threads, smem, st>>>(plan, d_tracks + t0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace thb
