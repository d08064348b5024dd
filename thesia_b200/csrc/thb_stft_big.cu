// thb_stft_big.cu -- K1/K2/K3 for n_fft == 16384 (BASELINE config 4: win 16384, hop 1024, 96 kHz): the large-FFT
// shared-memory path.  One CTA of 256 threads transforms TWO consecutive frames at once in packed f32x2 arithmetic
// (thb_packed.cuh: .x = frame A, .y = frame B).
//
// The frame is packed as 8192 complex z[m] = x[2m] + i x[2m+1]; 8192 = 32 x 16 x 16:
//   step 1  thread t holds z[256 n1 + t], n1 = 0..31 (coalesced 8-byte loads straight from the PCM, window folded in):
//           32-point DFT in registers, twiddle W_512^(n2 k1), 16-byte stores to shared memory
//   step 2  thread (k1, n3) x 2: 16-point DFT over n2, twiddle W_8192^(n3 (k1 + 32 k2)) as the product of two small
//           shared-memory tables, stored back in place
//   step 3  thread (k1, k2) x 2: 16-point DFT over n3 -> Z[k1 + 32 k2 + 512 k3]; after a barrier the spectrum is laid
//           out in natural order (one pad element every 32) so that the real split reads Z[k] and Z[8192 - k]
//           conflict-free and the dB rows leave as coalesced stores
//   split   X[k], X[8192 - k] from Z[k], Z[8192 - k] (k = 0 pairs with itself and yields bins 0 and 8192),
//           |X|^2 -> dB (linear) or |X| -> back into the FFT buffer, in place (mel)
//   mel     band-major pieces of <= 32 bins, 32 pieces per warp in lock step (coalesced step-major weights), partial
//           sums added per band in ascending-bin order
// Every 16-byte shared-memory element is {re.A, re.B, im.A, im.B}; rows of 16 elements are padded to 17, which keeps
// the 128-bit accesses of all three steps on distinct bank groups.  Every frame of a channel goes through this one
// kernel (file edges take a reflect-indexed load, utils.rs:111-137, an odd last frame is paired with itself), so a
// frame's result does not depend on how a file is sharded.
//
//   perform_stft (stft.rs:16-149), Complex::norm (spectrogram.rs:200), linspec.dot(mel_fb) (spectrogram.rs:207),
//   dB_from_amp (decibel.rs:198-202), find_min_max (mod.rs:169-178).
#include "thb_packed.cuh"

#include <cstdlib>
#include <type_traits>

namespace thb {

namespace {

using namespace packed;

constexpr int kCols = 256;       // columns of step 1: z[256 n1 + col]
constexpr int kRow17 = 17;       // padded row of 16 elements
constexpr int kPairsPerItem = 8; // frame pairs a CTA takes at a time
// element index of A[k1][x][y] (steps 1-3) and of Z[k] (natural order)
__device__ __forceinline__ int idx3(int k1, int x, int y) { return k1 * (16 * kRow17) + x * kRow17 + y; }
// natural order: one pad element every P = min(R1, 32) elements (shift = log2 P).  Step 3 writes Z[k1 + R1 k2 + ...] with
// k2 running over the lanes: a stride of R1 elements of 16 bytes would put a quarter warp on one or two bank groups;
// with the pad the stride becomes R1 + 1 (odd) and the eight 16-byte stores of a quarter warp hit eight bank groups.
constexpr int pad_shift(int r1) { return r1 == 16 ? 4 : (r1 == 8 ? 3 : 5); }  // R1 = 2 keeps P = 32: a stride of 2 is a 2-way conflict at worst
// R1 = n_fft / 512 = 8, 16 or 32: the frame is 256 R1 complex points = R1 x 16 x 16
// the three-step layout needs 272 R1 elements, the natural-order one NC (1 + 1/P) (+ the 31 bins a mel piece may read past its end)
constexpr int natural_elems(int r1) { return 256 * r1 + ((256 * r1) >> pad_shift(r1)) + 40; }
constexpr int buf_elems(int r1) { return r1 * 16 * kRow17 > natural_elems(r1) ? r1 * 16 * kRow17 : natural_elems(r1); }  // R1 = 32: 136 KB

// in-register R-point DFT of both frames; output X[k] is left in v[perm_r<R>(k)]
template <int R>
__device__ __forceinline__ void dft_r(cx (&v)[R]) {
    if constexpr (R == 32) dft32p(v);
    else if constexpr (R == 16) dft16p(v);
    else if constexpr (R == 8) dft8p(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    else {
        const cx a = v[0], b = v[1];
        v[0] = cadd2(a, b);
        v[1] = csub2(a, b);
    }
}
template <int R>
__device__ __forceinline__ constexpr int perm_r(int k) { return R == 32 ? perm32(k) : (R == 16 ? perm16(k) : k); }

struct alignas(16) Elem {
    float2 re, im;
};

__device__ __forceinline__ Elem to_elem(const cx &c) { return Elem{c.re, c.im}; }
__device__ __forceinline__ cx to_cx(const Elem &e) { return cx{e.re, e.im}; }

__device__ __forceinline__ float mag_of(float s, float re, float im) {
    return (s > kPowTiny && s < kPowHuge) ? sqrt_ftz(s) : hypotf(re, im);
}
// 10 log10(re^2 + im^2) = 20 log10 |X|; the square is only trusted inside [2^-80, 2^100]
__device__ __forceinline__ float db_of(float s, float re, float im) {
    if (s > kPowTiny && s < kPowHuge) return kDbPerLog2Pow * lg2_ftz(s);
    return amp_to_db(hypotf(re, im));
}

// kThreads = 256: one thread per column, 32-point DFT in registers (dft32p).
// kThreads = 512: two threads per column and per (k1, n3) / (k1, k2) task.  Thread h = t >> 8 of a column computes the
// outputs k1 = 2 q + h of the 32-point DFT as one radix-2 stage + a 16-point DFT (y_0[n] = x[n] + x[n+16],
// y_1[n] = (x[n] - x[n+16]) W_32^n; X[2q + h] = DFT16(y_h)[q]), and steps 2 / 3 take one task per thread instead of
// two: the same shared-memory layout and tables, 16 warps per SM instead of 8, <= 128 registers per thread.
// WS: the padded window (64 KB) lives in shared memory instead of L2 / L1
// R1 = 8 / 16 (n_fft 4096 / 8192): the same three steps with an 8- / 16-point first DFT, one task per thread
// (kThreads = 16 R1 = 128 / 256; R1 = 8 walks two columns per thread in step 1), 35 / 70 KB of shared memory and
// four / two CTAs per SM.  R1 = 2 (n_fft 1024): one warp per CTA walks eight columns in step 1, 11 KB of shared
// memory, sixteen CTAs per SM.
template <bool MEL, int kThreads, bool WS, int R1>
__global__ void __launch_bounds__(kThreads, R1 == 2 ? 16 : (R1 == 8 ? 4 : (R1 == 16 ? 2 : 1))) stft16384_kernel(const PlanDev p, const TrackDesc *__restrict__ tracks,
                                                               long long n_items, long long items_per_track) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kNC = 256 * R1;             // complex points
    constexpr int kBufElems = buf_elems(R1);
    constexpr int kPadShift = pad_shift(R1);
    auto idxn = [](int k) { return k + (k >> kPadShift); };
    static_assert(kThreads == 16 * R1 || (R1 == 32 && kThreads == 256), "one (k1, n3) task per thread, or the two-task R1 = 32 variant");
    Elem *buf = reinterpret_cast<Elem *>(smem_raw);                         // [kBufElems]
    float2 *tw_a = reinterpret_cast<float2 *>(buf + kBufElems);   // [R1 - 1][16] W_(16 R1)^(n2 k1)
    float2 *tw_b = tw_a + (R1 - 1) * 16;                                    // [R1][16] W_NC^(n3 k1)
    float2 *tw_c = tw_b + R1 * 16;                                          // [16][16] W_256^(n3 k2)
    float *wsm = reinterpret_cast<float *>(tw_c + 16 * 16);                 // [n_fft] padded window (WS only)
    __shared__ float red_max[kThreads / 32], red_nmin[kThreads / 32];
    // After the real split (and a barrier) the FFT buffer is reused: |X[k]| of both frames goes to float2 slot k, the mel
    // partial sums follow at slot NC + 64.
    // compact |X[k]| of both frames at float2 slot kMagBase + k (the bin-major walk may start up to 15 bins before bin 0 and
    // reach past bin NC on zero weights), then the mel partial sums
    float2 *cmag = reinterpret_cast<float2 *>(buf) + k2048::kMagBase;
    const bool bin_major = MEL && p.mi_blob != nullptr;
    float2 *cpart = cmag + (bin_major ? ((p.mi_max_reach + 2) & ~1) : kNC + 64);
    auto mag_at = [&](int k) -> float2 & { return cmag[k]; };
    auto part_at = [&](int q) -> float2 & { return cpart[q]; };

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int col = t & (kCols - 1), hsel = t >> 8;  // kThreads == 512: hsel picks the parity of k1 / the upper k1 half
    for (int i = t; i < (R1 - 1) * 16 + R1 * 16 + 16 * 16; i += kThreads) tw_a[i] = __ldg(&p.big_tw[i]);
    if constexpr (WS) {
        for (int i = t; i < 2 * kNC / 4; i += kThreads)
            reinterpret_cast<float4 *>(wsm)[i] = __ldg(reinterpret_cast<const float4 *>(p.big_wpad) + i);
    }
    // (the mel walks read a few bins before bin 0 / past bin NC with weight 0: those slots are inside the buffer, and what
    // they find there is left-over FFT data, finite whenever the frame is)
    __syncthreads();
    const int half = p.win / 2;
    const float2 w_own = __ldg(&p.twiddle[t]);  // W_n_fft^t (real split, 512-thread variant)

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long track = item / items_per_track, chunk = item - track * items_per_track;
        const TrackDesc d = tracks[track];
        const long long f_begin = chunk * (2 * kPairsPerItem);
        if (f_begin >= d.n_frames) continue;
        const long long f_end = min(f_begin + 2 * kPairsPerItem, d.n_frames);
        float lmax = -CUDART_INF_F, lnmin = -CUDART_INF_F;

        for (long long fa = f_begin; fa < f_end; fa += 2) {
            const bool has_b = fa + 1 < f_end;
            const long long fb = has_b ? fa + 1 : fa;  // an odd last frame is computed twice, stored once
            // ---- step 1: load + window, 32-point DFT over n1 ----
            if constexpr (kThreads != 512) {
#pragma unroll 1
            for (int c = t; c < kCols; c += kThreads) {   // one column per thread (two when R1 = 8)
            cx v[R1];
            {
                const long long tap0[2] = {(d.frame_begin + fa) * p.hop - half, (d.frame_begin + fb) * p.hop - half};
                const float *wsrc = p.big_wpad + 2 * c;
                bool interior[2], inside[2];   // inside: plain f32 loads, no reflection; interior: and 8-byte aligned
#pragma unroll
                for (int f = 0; f < 2; f++) {
                    const long long first = tap0[f] - p.pad_left;  // file index of FFT position 0
                    inside[f] = !d.pcm_i16 && tap0[f] >= 0 && tap0[f] + p.win <= d.full_len && first >= d.pcm_offset &&
                                first + 2 * kNC <= d.pcm_offset + d.slice_len;
                    interior[f] = inside[f] && ((reinterpret_cast<uintptr_t>(d.pcm + (first - d.pcm_offset))) & 7) == 0;
                }
                bool done = false;
                if constexpr (R1 == 32) {
                    if (interior[0] && interior[1] && has_b && p.hop == 1024) {
                        // frame B starts hop = 1024 samples = 2 rows of 256 complex later: 34 loads serve both frames
                        const float *src_a = d.pcm + (tap0[0] - p.pad_left - d.pcm_offset) + 2 * c;
                        float2 r[34];
#pragma unroll
                        for (int n1 = 0; n1 < 34; n1++) r[n1] = __ldg(reinterpret_cast<const float2 *>(src_a + 512 * n1));
#pragma unroll
                        for (int n1 = 0; n1 < 32; n1++) {
                            const float2 w = __ldg(reinterpret_cast<const float2 *>(wsrc + 512 * n1));
                            v[n1].re = make_float2(r[n1].x * w.x, r[n1 + 2].x * w.x);
                            v[n1].im = make_float2(r[n1].y * w.y, r[n1 + 2].y * w.y);
                        }
                        done = true;
                    }
                }
                if (done) {
                } else if (interior[0] && interior[1]) {
                    const float *src_a = d.pcm + (tap0[0] - p.pad_left - d.pcm_offset) + 2 * c;
                    const float *src_b = d.pcm + (tap0[1] - p.pad_left - d.pcm_offset) + 2 * c;
#pragma unroll
                    for (int n1 = 0; n1 < R1; n1++) {
                        const float2 xa = __ldg(reinterpret_cast<const float2 *>(src_a + 512 * n1));
                        const float2 xb = __ldg(reinterpret_cast<const float2 *>(src_b + 512 * n1));
                        const float2 w = __ldg(reinterpret_cast<const float2 *>(wsrc + 512 * n1));
                        v[n1].re = make_float2(xa.x * w.x, xb.x * w.x);
                        v[n1].im = make_float2(xa.y * w.y, xb.y * w.y);
                    }
                } else if (inside[0] && inside[1]) {
                    // an odd hop or an odd channel start: the same samples through 4-byte loads
                    const float *src_a = d.pcm + (tap0[0] - p.pad_left - d.pcm_offset) + 2 * c;
                    const float *src_b = d.pcm + (tap0[1] - p.pad_left - d.pcm_offset) + 2 * c;
#pragma unroll
                    for (int n1 = 0; n1 < R1; n1++) {
                        const float2 w = __ldg(reinterpret_cast<const float2 *>(wsrc + 512 * n1));
                        v[n1].re = make_float2(__ldg(src_a + 512 * n1) * w.x, __ldg(src_b + 512 * n1) * w.x);
                        v[n1].im = make_float2(__ldg(src_a + 512 * n1 + 1) * w.y, __ldg(src_b + 512 * n1 + 1) * w.y);
                    }
                } else {
                    // file edges (numpy-style reflect, utils.rs:111-137), 16-bit PCM
                    auto tap = [&](int f, int pos) -> float {
                        const int a = pos - p.pad_left;
                        if (a < 0 || a >= p.win) return 0.0f;
                        return pcm_sample(d, reflect_index(tap0[f] + a, d.full_len) - d.pcm_offset);
                    };
#pragma unroll
                    for (int n1 = 0; n1 < R1; n1++) {
                        const int pos = 512 * n1 + 2 * c;
                        const float2 w = __ldg(reinterpret_cast<const float2 *>(wsrc + 512 * n1));
                        v[n1].re = make_float2(tap(0, pos) * w.x, tap(1, pos) * w.x);
                        v[n1].im = make_float2(tap(0, pos + 1) * w.y, tap(1, pos + 1) * w.y);
                    }
                }
            }
            dft_r<R1>(v);
            {
                const int n2 = c >> 4, n3 = c & 15;
#pragma unroll
                for (int k1 = 0; k1 < R1; k1++) {
                    cx o = v[perm_r<R1>(k1)];
                    if (k1) {
                        const float2 w = tw_a[(k1 - 1) * 16 + n2];
                        o = cmul_s(o, w.x, w.y);
                    }
                    buf[idx3(k1, n2, n3)] = to_elem(o);
                }
            }
            }
            } else {
                // two threads per column: this one computes the outputs k1 = 2 q + hsel
                const long long tap0[2] = {(d.frame_begin + fa) * p.hop - half, (d.frame_begin + fb) * p.hop - half};
                const float *wsrc = (WS ? wsm : p.big_wpad) + 2 * col;
                bool interior[2];
#pragma unroll
                for (int f = 0; f < 2; f++) {
                    const long long first = tap0[f] - p.pad_left;
                    interior[f] = !d.pcm_i16 && tap0[f] >= 0 && tap0[f] + p.win <= d.full_len && first >= d.pcm_offset &&
                                  first + 2 * kNC <= d.pcm_offset + d.slice_len &&
                                  ((reinterpret_cast<uintptr_t>(d.pcm + (first - d.pcm_offset))) & 7) == 0;
                }
                cx y[16];
                auto fold = [&](int n, float2 xa0, float2 xb0, float2 xa1, float2 xb1) {
                    // windowed samples of rows n and n + 16 of both frames, then the radix-2 stage
                    const float2 w0 = WS ? *reinterpret_cast<const float2 *>(wsrc + 512 * n) : __ldg(reinterpret_cast<const float2 *>(wsrc + 512 * n));
                    const float2 w1 = WS ? *reinterpret_cast<const float2 *>(wsrc + 512 * (n + 16))
                                         : __ldg(reinterpret_cast<const float2 *>(wsrc + 512 * (n + 16)));
                    cx lo, hi;
                    lo.re = make_float2(xa0.x * w0.x, xb0.x * w0.x);
                    lo.im = make_float2(xa0.y * w0.y, xb0.y * w0.y);
                    hi.re = make_float2(xa1.x * w1.x, xb1.x * w1.x);
                    hi.im = make_float2(xa1.y * w1.y, xb1.y * w1.y);
                    y[n] = hsel ? csub2(lo, hi) : cadd2(lo, hi);
                };
                if (interior[0] && interior[1] && has_b && p.hop == 1024) {
                    // frame B is frame A two rows later: rows 0 .. 33 serve both frames (34 distinct loads)
                    const float *src_a = d.pcm + (tap0[0] - p.pad_left - d.pcm_offset) + 2 * col;
                    float2 r[34];
#pragma unroll
                    for (int n1 = 0; n1 < 34; n1++) r[n1] = __ldg(reinterpret_cast<const float2 *>(src_a + 512 * n1));
#pragma unroll
                    for (int n = 0; n < 16; n++) fold(n, r[n], r[n + 2], r[n + 16], r[n + 18]);
                } else if (interior[0] && interior[1]) {
                    const float *src_a = d.pcm + (tap0[0] - p.pad_left - d.pcm_offset) + 2 * col;
                    const float *src_b = d.pcm + (tap0[1] - p.pad_left - d.pcm_offset) + 2 * col;
#pragma unroll
                    for (int n = 0; n < 16; n++)
                        fold(n, __ldg(reinterpret_cast<const float2 *>(src_a + 512 * n)),
                             __ldg(reinterpret_cast<const float2 *>(src_b + 512 * n)),
                             __ldg(reinterpret_cast<const float2 *>(src_a + 512 * (n + 16))),
                             __ldg(reinterpret_cast<const float2 *>(src_b + 512 * (n + 16))));
                } else {
                    auto tap = [&](int f, int pos) -> float {
                        const int a = pos - p.pad_left;
                        if (a < 0 || a >= p.win) return 0.0f;
                        return pcm_sample(d, reflect_index(tap0[f] + a, d.full_len) - d.pcm_offset);
                    };
#pragma unroll
                    for (int n = 0; n < 16; n++) {
                        const int p0 = 512 * n + 2 * col, p1 = 512 * (n + 16) + 2 * col;
                        fold(n, make_float2(tap(0, p0), tap(0, p0 + 1)), make_float2(tap(1, p0), tap(1, p0 + 1)),
                             make_float2(tap(0, p1), tap(0, p1 + 1)), make_float2(tap(1, p1), tap(1, p1 + 1)));
                    }
                }
                if (hsel) {
#pragma unroll
                    for (int n = 1; n < 16; n++) y[n] = cmul_s(y[n], kC32[n], -kS32[n]);   // * W_32^n
                }
                dft16p(y);
                const int n2 = col >> 4, n3 = col & 15;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int k1 = 2 * q + hsel;
                    cx o = y[perm16(q)];
                    if (k1) {
                        const float2 w = tw_a[(k1 - 1) * 16 + n2];
                        o = cmul_s(o, w.x, w.y);
                    }
                    buf[idx3(k1, n2, n3)] = to_elem(o);
                }
            }
            __syncthreads();
            // ---- step 2: 16-point DFT over n2, twiddle W_NC^(n3 (k1 + R1 k2)) ----
#pragma unroll 1
            for (int task = t; task < 16 * R1; task += kThreads) {
                const int k1 = task >> 4, n3 = task & 15;
                cx u[16];
#pragma unroll
                for (int n2 = 0; n2 < 16; n2++) u[n2] = to_cx(buf[idx3(k1, n2, n3)]);
                dft16p(u);
                // W_8192^(n3 (k1 + 32 k2)) = W_8192^(n3 k1) * W_256^(n3 k2): two small shared-memory tables instead of
                // a 64 KB one in L2 (the product is computed once for both frames)
                const float2 wb = tw_b[k1 * 16 + n3];
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++) {
                    const float2 wc = tw_c[k2 * 16 + n3];
                    const float c = fmaf(-wb.y, wc.y, wb.x * wc.x), sn = fmaf(wb.x, wc.y, wb.y * wc.x);
                    buf[idx3(k1, k2, n3)] = to_elem(cmul_s(u[perm16(k2)], c, sn));
                }
            }
            __syncthreads();
            // ---- step 3: 16-point DFT over n3 -> Z[k1 + R1 k2 + 16 R1 k3], then natural order ----
            if constexpr (kThreads == 256 && R1 == 32) {
                cx u0[16], u1[16];
                const int k2 = t & 15, k1a = t >> 4, k1b = k1a + 16;
#pragma unroll
                for (int n3 = 0; n3 < 16; n3++) {
                    u0[n3] = to_cx(buf[idx3(k1a, k2, n3)]);
                    u1[n3] = to_cx(buf[idx3(k1b, k2, n3)]);
                }
                dft16p(u0);
                dft16p(u1);
                __syncthreads();
#pragma unroll
                for (int k3 = 0; k3 < 16; k3++) {
                    buf[idxn(k1a + 32 * k2 + 512 * k3)] = to_elem(u0[perm16(k3)]);
                    buf[idxn(k1b + 32 * k2 + 512 * k3)] = to_elem(u1[perm16(k3)]);
                }
            } else {
                cx u[16];
                const int k2 = t & 15, k1 = t >> 4;   // kThreads == 16 R1: one (k1, k2) task per thread
#pragma unroll
                for (int n3 = 0; n3 < 16; n3++) u[n3] = to_cx(buf[idx3(k1, k2, n3)]);
                dft16p(u);
                __syncthreads();
#pragma unroll
                for (int k3 = 0; k3 < 16; k3++) buf[idxn(k1 + R1 * k2 + 16 * R1 * k3)] = to_elem(u[perm16(k3)]);
            }
            __syncthreads();
            // ---- real split: pairs (k, 8192 - k), k = t + 256 j; |X|^2 -> dB, or |X| -> mag[] ----
            float *orow_a = d.out + fa * p.n_bins, *orow_b = d.out + fb * p.n_bins;
            constexpr int kJ = kNC / 2 / kThreads;   // k = t + kThreads j, j = 0 .. kJ (k = NC / 2 pairs with itself: thread 0 alone)
            auto split = [&](int j, float2 &mag_lo, float2 &mag_hi) {
                const int k = t + kThreads * j;
                if (j == kJ && t != 0) return;
                const int kp = (kNC - k) & (kNC - 1);
                const cx zk = to_cx(buf[idxn(k)]), zn = to_cx(buf[idxn(kp)]);
                const f2 er = padd(zk.re, zn.re), ei = psub(zk.im, zn.im), dr = psub(zk.re, zn.re), di = padd(zk.im, zn.im);
                float2 w;
                if constexpr (kThreads == 512) {
                    // W_16384^(t + 512 j) = W_16384^t * W_32^j (R1 = 32): the thread's own factor sits in a register pair, the
                    // other is a 9-entry constant table -- no L2 round trip per bin pair
                    const float c = kC32[j], sn = -kS32[j];
                    w = make_float2(fmaf(-w_own.y, sn, w_own.x * c), fmaf(w_own.x, sn, w_own.y * c));
                } else {
                    w = __ldg(&p.twiddle[k]);
                }
                const f2 wr = pfma(di, bc(-w.y), pmul(dr, bc(w.x))), wi = pfma(dr, bc(w.y), pmul(di, bc(w.x)));
                const f2 ar = padd(er, wi), ai = psub(ei, wr), br = psub(er, wi), bi = padd(ei, wr);
                const f2 sa = pfma(ar, ar, pmul(ai, ai)), sb = pfma(br, br, pmul(bi, bi));
                const int k_hi = kNC - k;  // bin of the partner: NC for k = 0, NC / 2 for k = NC / 2 (same bin)
                if (MEL) {
                    mag_lo = make_float2(mag_of(sa.x, ar.x, ai.x), mag_of(sa.y, ar.y, ai.y));
                    mag_hi = make_float2(mag_of(sb.x, br.x, bi.x), mag_of(sb.y, br.y, bi.y));
                } else {
                    const float a0 = db_of(sa.x, ar.x, ai.x), a1 = db_of(sa.y, ar.y, ai.y);
                    orow_a[k] = a0;
                    lmax = fmaxf(lmax, a0);
                    lnmin = fmaxf(lnmin, -a0);
                    if (has_b) {
                        orow_b[k] = a1;
                        lmax = fmaxf(lmax, a1);
                        lnmin = fmaxf(lnmin, -a1);
                    }
                    if (k_hi != k) {
                        const float b0 = db_of(sb.x, br.x, bi.x), b1 = db_of(sb.y, br.y, bi.y);
                        orow_a[k_hi] = b0;
                        lmax = fmaxf(lmax, b0);
                        lnmin = fmaxf(lnmin, -b0);
                        if (has_b) {
                            orow_b[k_hi] = b1;
                            lmax = fmaxf(lmax, b1);
                            lnmin = fmaxf(lnmin, -b1);
                        }
                    }
                }
            };
            if constexpr (MEL) {
                // The magnitudes wait in registers until every thread has read its Z[k], Z[NC - k]; then they are stored
                // COMPACTLY (8 bytes per bin, bin k at float2 slot k of the buffer) so that the mel walk's 32 lanes, each on
                // its own run of bins, spread over sixteen 8-byte bank pairs instead of the eight a 16-byte element stride
                // would leave them.
                float2 mg[2 * (kJ + 1)];
#pragma unroll
                for (int j = 0; j <= kJ; j++) split(j, mg[2 * j], mg[2 * j + 1]);
                __syncthreads();
#pragma unroll
                for (int j = 0; j <= kJ; j++) {
                    const int k = t + kThreads * j;
                    if (j == kJ && t != 0) continue;
                    cmag[k] = mg[2 * j];
                    if (kNC - k != k) cmag[kNC - k] = mg[2 * j + 1];
                }
            } else {
                float2 unused0, unused1;
#pragma unroll 4
                for (int j = 0; j <= kJ; j++) split(j, unused0, unused1);
            }
            __syncthreads();
            if (MEL && bin_major) {
                // ---- sparse mel, bin-major (the schedule of the n_fft 2048 kernels, thb_host.hpp MelItems, read from global
                // memory): every |X[k]| is read once and multiplied by a (rise, fall) weight pair; the lanes of a half warp
                // start on sixteen different residues mod 16, so the float2 reads are conflict-free; the warps share the groups
                const k2048::MelView mv(p.mi_blob);
                {
                    // mel_walk (thb_stft2048.cuh) with the schedule in L2 instead of shared memory: the group headers of
                    // this warp are fetched together, and all weights of a group are in registers before the first use,
                    // so a group costs one L2 round trip, not one per step
                    constexpr int kNW = kThreads / 32, kMaxT2 = 9;   // T <= kMelPieceMax + 3, even
                    constexpr int kBatch = 4;                        // group headers fetched together
                    const int n_slots = mv.n_groups * 32;
                    for (int gb = warp; gb < mv.n_groups; gb += kBatch * kNW) {
                        uint2 gh[kBatch];
                        int st[kBatch];
#pragma unroll
                        for (int b = 0; b < kBatch; b++) {
                            const int g = gb + b * kNW;
                            gh[b] = g < mv.n_groups ? __ldg(mv.grp + g) : make_uint2(0u, 0u);
                            st[b] = g < mv.n_groups ? __ldg(mv.start + g * 32 + lane) : 0;
                        }
#pragma unroll
                        for (int b = 0; b < kBatch; b++) {
                            const int g = gb + b * kNW;
                            if (g >= mv.n_groups) break;
                            const float4 *wq = reinterpret_cast<const float4 *>(mv.base + gh[b].y) + lane;
                            const int T2 = static_cast<int>(gh[b].x) >> 1;
                            float4 wv[kMaxT2];
#pragma unroll
                            for (int i = 0; i < kMaxT2; i++) wv[i] = i < T2 ? __ldg(wq + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float2 *mq = cmag + st[b];
                            f2 rise = make_float2(0.f, 0.f), fall = rise;
#pragma unroll
                            for (int i = 0; i < kMaxT2; i++) {
                                if (i < T2) {
                                    const float2 m0 = mq[2 * i], m1 = mq[2 * i + 1];
                                    rise = pfma(m0, bc(wv[i].x), rise);
                                    fall = pfma(m0, bc(wv[i].y), fall);
                                    rise = pfma(m1, bc(wv[i].z), rise);
                                    fall = pfma(m1, bc(wv[i].w), fall);
                                }
                            }
                            cpart[g * 32 + lane] = rise;
                            cpart[n_slots + g * 32 + lane] = fall;
                        }
                    }
                    if (t == 0) cpart[2 * n_slots] = make_float2(0.f, 0.f);   // the padding slot of the gather
                }
                __syncthreads();
                for (int r = warp; 32 * r < mv.n_mel; r += kThreads / 32) {
                    const int m = 32 * r + lane;
                    // mel_band (thb_stft2048.cuh) with the slot ids in L2: eight ids are fetched together, then summed in
                    // the schedule's order
                    const uint2 rd = __ldg(mv.rounds + r);
                    const uint16_t *row = mv.goff + rd.y * 32 + lane;
                    f2 acc = make_float2(0.f, 0.f);
                    for (uint32_t j0 = 0; j0 < rd.x; j0 += 8) {
                        uint16_t id[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) id[j] = j0 + j < rd.x ? __ldg(row + 32 * (j0 + j)) : static_cast<uint16_t>(2 * mv.n_groups * 32);
#pragma unroll
                        for (int j = 0; j < 8; j++) acc = padd(acc, cpart[id[j]]);   // the padding slot holds zero
                    }
                    if (m >= mv.n_mel) continue;
                    const float a0 = kDbPerLog2Amp * lg2_ftz(acc.x), a1 = kDbPerLog2Amp * lg2_ftz(acc.y);
                    orow_a[m] = a0;
                    lmax = fmaxf(lmax, a0);
                    lnmin = fmaxf(lnmin, -a0);
                    if (has_b) {
                        orow_b[m] = a1;
                        lmax = fmaxf(lmax, a1);
                        lnmin = fmaxf(lnmin, -a1);
                    }
                }
                __syncthreads();
            } else if (MEL) {
                // ---- sparse mel, band-major fallback (banks the bin-major schedule does not cover): a warp walks 32 pieces
                // (<= 31 consecutive bins of one band each) in lock step ----
                const int n_groups = (p.big_n_pieces + 31) >> 5;
                for (int g = warp; g < n_groups; g += kThreads / 32) {
                    const int q = 32 * g + lane;
                    const uint2 gh = __ldg(reinterpret_cast<const uint2 *>(p.big_pieces + 32 * n_groups) + g);  // {steps, offset}
                    const int k0 = static_cast<int>(__ldg(&p.big_pieces[q]));
                    const float *w = p.big_w + gh.y + lane;
                    // the weight table (150 KB for 1621 bands) lives in L2: all loads of a group are issued before the
                    // first use, so a group costs one L2 round trip, not one per step
                    f2 acc = make_float2(0.0f, 0.0f);
                    const int T = static_cast<int>(gh.x);
                    auto walk = [&](auto nmax) {  // unrolled for groups of up to nmax steps (padding lanes: weight 0)
                        constexpr int N = decltype(nmax)::value;
                        float wv[N];
#pragma unroll
                        for (int i = 0; i < N; i++) wv[i] = i < T ? __ldg(w + 32 * i) : 0.0f;
#pragma unroll
                        for (int i = 0; i < N; i++)
                            if (i < T) acc = pfma(mag_at(k0 + i), bc(wv[i]), acc);
                    };
                    if (T <= 8) walk(std::integral_constant<int, 8>{});
                    else if (T <= 16) walk(std::integral_constant<int, 16>{});
                    else walk(std::integral_constant<int, 32>{});
                    if (q < p.big_n_pieces) part_at(q) = acc;
                }
                __syncthreads();
                for (int m = t; m < p.n_mel; m += kThreads) {
                    const uint32_t q0 = __ldg(&p.big_pptr[m]), q1 = __ldg(&p.big_pptr[m + 1]);
                    f2 acc = make_float2(0.0f, 0.0f);
                    for (uint32_t q = q0; q < q1; q++) acc = padd(acc, part_at(q));
                    const float a0 = kDbPerLog2Amp * lg2_ftz(acc.x), a1 = kDbPerLog2Amp * lg2_ftz(acc.y);
                    orow_a[m] = a0;
                    lmax = fmaxf(lmax, a0);
                    lnmin = fmaxf(lnmin, -a0);
                    if (has_b) {
                        orow_b[m] = a1;
                        lmax = fmaxf(lmax, a1);
                        lnmin = fmaxf(lnmin, -a1);
                    }
                }
                __syncthreads();
            }
        }
        // ---- per-channel {max, -min}: warp shuffle -> shared -> one atomic pair per CTA and item ----
        lmax = warp_max(lmax);
        lnmin = warp_max(lnmin);
        if (lane == 0) {
            red_max[warp] = lmax;
            red_nmin[warp] = lnmin;
        }
        __syncthreads();
        if (warp == 0) {
            float a = lane < kThreads / 32 ? red_max[lane] : -CUDART_INF_F;
            float b = lane < kThreads / 32 ? red_nmin[lane] : -CUDART_INF_F;
            a = warp_max(a);
            b = warp_max(b);
            if (lane == 0) {
                atomic_max_float(&d.minmax[0], a);
                atomic_max_float(&d.minmax[1], b);
            }
        }
        __syncthreads();
    }
}

int big_r1(const PlanDev &p) { return p.n_fft / 512; }
size_t big_smem_bytes(const PlanDev &p, bool ws = false) {
    const int r1 = big_r1(p);
    return sizeof(Elem) * buf_elems(r1) + sizeof(float2) * ((r1 - 1) * 16 + r1 * 16 + 16 * 16) + (ws ? sizeof(float) * p.n_fft : 0);
}

}  // namespace

int stft_big_buffer_slots(int n_fft) { return 2 * buf_elems(n_fft / 512); }

bool stft_big_supported(const PlanDev &p) {
    if ((p.n_fft != 1024 && p.n_fft != 4096 && p.n_fft != 8192 && p.n_fft != 16384) || !p.big_wpad || !p.big_tw) return false;
    if (p.n_mel && (!p.big_pieces || !p.big_pptr || !p.big_w || p.big_n_pieces > 2 * buf_elems(big_r1(p)) - (p.n_fft / 2 + 64 + k2048::kMagBase)))
        return false;   // compact magnitudes + one partial sum per piece share the FFT buffer
    if (p.n_mel && p.mi_blob &&
        (p.mi_min_start < -(k2048::kMagBase - 1) ||
         k2048::kMagBase + ((p.mi_max_reach + 2) & ~1) + 2 * p.mi_groups * 32 + 1 > 2 * buf_elems(big_r1(p))))
        return false;   // (the plan builder only keeps a bin-major schedule that fits: see thb_api.cu)
    return big_smem_bytes(p, true) <= 226 * 1024 || big_smem_bytes(p) <= 226 * 1024;
}

namespace {
template <bool MEL, int NT, bool WS, int R1>
cudaError_t launch_big_nt(const PlanDev &plan, const TrackDesc *d_tracks, long long n_items, long long items_per_track, int sm_count,
                          cudaStream_t st) {
    const size_t smem = big_smem_bytes(plan, WS);
    cudaError_t e = cudaFuncSetAttribute(stft16384_kernel<MEL, NT, WS, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const long long slots = static_cast<long long>(sm_count) * (R1 == 2 ? 16 : (R1 == 8 ? 4 : (R1 == 16 ? 2 : 1)));  // persistent grid
    const int grid = static_cast<int>(n_items < slots ? n_items : slots);
    stft16384_kernel<MEL, NT, WS, R1><<<grid, NT, smem, st>>>(plan, d_tracks, n_items, items_per_track);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_stft_big(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, long long max_frames,
                            int sm_count, cudaStream_t st) {
    if (n_tracks <= 0 || max_frames <= 0) return cudaSuccess;
    const long long items_per_track = (max_frames + 2 * kPairsPerItem - 1) / (2 * kPairsPerItem);
    const long long n_items = items_per_track * n_tracks;
    const bool mel = plan.n_mel != 0;
    if (plan.n_fft == 1024)
        return mel ? launch_big_nt<true, 32, false, 2>(plan, d_tracks, n_items, items_per_track, sm_count, st)
                   : launch_big_nt<false, 32, false, 2>(plan, d_tracks, n_items, items_per_track, sm_count, st);
    if (plan.n_fft == 4096)
        return mel ? launch_big_nt<true, 128, false, 8>(plan, d_tracks, n_items, items_per_track, sm_count, st)
                   : launch_big_nt<false, 128, false, 8>(plan, d_tracks, n_items, items_per_track, sm_count, st);
    if (plan.n_fft == 8192)
        return mel ? launch_big_nt<true, 256, false, 16>(plan, d_tracks, n_items, items_per_track, sm_count, st)
                   : launch_big_nt<false, 256, false, 16>(plan, d_tracks, n_items, items_per_track, sm_count, st);
    // n_fft 16384.  A/B knobs: THB_BIG_THREADS = 256 | 512, THB_BIG_WSMEM = 0 | 1 (512 only)
    const char *e = getenv("THB_BIG_THREADS");
    const int nt = (e && atoi(e) == 256) ? 256 : 512;
    const char *w = getenv("THB_BIG_WSMEM");
    const bool ws = nt == 512 && !(w && atoi(w) == 0);
    if (mel) {
        if (nt == 256) return launch_big_nt<true, 256, false, 32>(plan, d_tracks, n_items, items_per_track, sm_count, st);
        if (ws) return launch_big_nt<true, 512, true, 32>(plan, d_tracks, n_items, items_per_track, sm_count, st);
        return launch_big_nt<true, 512, false, 32>(plan, d_tracks, n_items, items_per_track, sm_count, st);
    }
    if (nt == 256) return launch_big_nt<false, 256, false, 32>(plan, d_tracks, n_items, items_per_track, sm_count, st);
    if (ws) return launch_big_nt<false, 512, true, 32>(plan, d_tracks, n_items, items_per_track, sm_count, st);
    return launch_big_nt<false, 512, false, 32>(plan, d_tracks, n_items, items_per_track, sm_count, st);
}

}  // namespace thb
