// thb_stft_pair.cu -- K1/K2/K3 for n_fft == 2048, TWO frames per warp in packed f32x2 arithmetic.
//
// Same dataflow as the warp-per-frame kernel (thb_stft_fast.cu: 1024 = 32 x 32 register FFT of the
// even/odd-packed frame, one shared-memory transpose, real split over one shuffle pair, sparse mel,
// dB through MUFU.LG2), but every register pair holds the same quantity of two consecutive frames
// (x = frame 2p, y = frame 2p+1) and all arithmetic is issued as sm_100 FADD2 / FMUL2 / FFMA2.  The
// compile-time DFT twiddles are broadcast immediates, the table twiddles / window / mel weights are
// broadcast register operands: they are loaded ONCE per frame pair, so both the FP32 issue slots and the
// shared-memory wavefronts per frame are about halved against the scalar kernel.
//
//   perform_stft (stft.rs:16-149), Complex::norm (spectrogram.rs:200), linspec.dot(mel_fb)
//   (spectrogram.rs:207), dB_from_amp (decibel.rs:198-202), find_min_max (mod.rs:169-178).
#include "thb_packed.cuh"

#include <cstdlib>

namespace thb {

namespace {

using namespace k2048;
using namespace packed;

struct PairSmem {
    float *wpad;         // [2048]   0.5 * window, zero outside the taps
    float2 *tw1;         // [31][32] W_1024^(lane * k1), k1 = 1..31
    float2 *tw2;         // [16][32] W_2048^(k_own(j, lane))
    float2 *tiles;       // [NW][tile_f2]
    const uint32_t *ms;  // MelItems blob
};

template <int NW>
__device__ __forceinline__ PairSmem carve(unsigned char *raw, int tile_f2) {
    PairSmem s;
    s.wpad = reinterpret_cast<float *>(raw);
    s.tw1 = reinterpret_cast<float2 *>(s.wpad + 2048);
    s.tw2 = s.tw1 + 31 * 32;
    s.tiles = s.tw2 + 16 * 32;
    s.ms = reinterpret_cast<const uint32_t *>(s.tiles + NW * tile_f2);
    return s;
}

// NW = warps per CTA; one persistent CTA per SM, so NW also sets the register budget (65536 / (32 NW)).
// Work item = (descriptor, tile of 8 NW consecutive frames = 4 frame pairs per warp); CTAs take items round
// robin, warps never meet at a block barrier after the tables are loaded.
//
// Contract (the launcher guarantees it, thb_api.cu): every frame of every descriptor is interior (no reflect
// padding, the whole n_fft span inside the slice), 8-byte aligned, and n_frames is even.  Edge frames, odd
// leftovers and unaligned channels go through the scalar kernel.
// I16: the channels hold 16-bit PCM.  A 32-bit load brings a sample pair; each half is sign-extended and
// converted (I2F, exact), and the 2^-15 of "s / 32768" rides on the window table, so every product equals the f32
// channel's bit for bit.  (Not the 1.5 * 2^23 exponent-pasting trick: the compiler distributes the window over its
// subtraction, (x - M) w -> fma(x, w, -M w), which is no longer exact.)
// HS > 0: hop == 64 HS, so frame B's row n1 is frame A's row n1 + HS and 32 + HS loads serve both frames.
// HS < 0: frames may start on any 4-byte boundary (odd hop such as the 44.1 kHz default's 441, odd channel start):
// the same samples through 4-byte loads, everything after the loads is identical.
// L2 prefetch of a byte range of global memory (one instruction, no register destination, no completion to wait for)
__device__ __forceinline__ void prefetch_l2_bulk(const void *addr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
}

// flags: bit 0 = while a tile is computed, pull the PCM of this CTA's NEXT tile into L2 (the loads of a frame pair
// then hit L2 instead of HBM: the pair's 40 loads sit at the head of a long dependent chain)
// M4: the lean mel walk / gather (mel_walk4 / mel_band4; same results bit for bit)
// PL (HS > 0): the PCM of the warp's NEXT frame pair is prefetched into L1 before the mel stage of the current one
// (three instructions per pair), so that its 40 loads hit L1 instead of stalling the window multiply on L2 / HBM
// DIRECT: the plan carries the band-major mel schedule (compile-time: the kernel sits close to the instruction-cache
// cliff noted in DESIGN.md, each variant holds only the mel code it runs)
template <bool MEL, int NW, bool I16, int HS, bool M4, bool PL, bool DIRECT>
__global__ void __launch_bounds__(NW * 32, 1) stft2048_pair_kernel(const PlanDev p,
                                                                        const TrackDesc *__restrict__ tracks,
                                                                        long long n_items, RescueList rescue, int flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kTile = 8 * NW;

    const int tile_f2 = tile_elems(p);
    const PairSmem sm = carve<NW>(smem_raw, tile_f2);
    for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) {
        float4 w = __ldg(reinterpret_cast<const float4 *>(p.fast_wpad) + i);
        if (I16) {
            const float k = 3.0517578125e-05f;  // 2^-15, exact
            w = make_float4(w.x * k, w.y * k, w.z * k, w.w * k);
        }
        reinterpret_cast<float4 *>(sm.wpad)[i] = w;
    }
    for (int i = threadIdx.x; i < (31 * 32 + 16 * 32) / 2; i += blockDim.x)
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
    float2 *mag = tile + kMagBase;  // mag[k] = (|X_A[k]|, |X_B[k]|)
    float2 *part = tile + part_base(p);
    const int half = p.win / 2;
    const int partner = (32 - lane) & 31;
    const int lane32 = lane ? lane : 32;
    const long long tiles_per_track = rescue.tiles_per_track;

    // (Handing the items out through a global counter -- two block barriers per tile -- was measured in round 2: 14 % slower,
    // 1.227 against 1.076 ms at 450 k frames.  The warps of a CTA drift apart by design; every barrier makes them wait.)
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long track = item / tiles_per_track, tile_idx = item - track * tiles_per_track;
        const TrackDesc d = tracks[track];
        const long long f_begin = tile_idx * kTile;
        if (f_begin >= d.n_frames) continue;
        const long long f_end = min(f_begin + kTile, d.n_frames);
        float lmax = -CUDART_INF_F, lnmin = -CUDART_INF_F;
        bool flagged = false;
#ifdef THB_PAIR_EXPERIMENTS
        if ((flags & 1) && lane == 0) {
            const long long nxt = item + gridDim.x;
            if (nxt < n_items) {
                const long long t2 = nxt / tiles_per_track, i2 = nxt - t2 * tiles_per_track;
                const TrackDesc *d2 = tracks + t2;
                const long long nf2 = __ldg(&d2->n_frames), fb2 = i2 * kTile;
                if (fb2 < nf2) {
                    constexpr int esz = I16 ? 2 : 4;
                    const long long fe2 = min(fb2 + kTile, nf2);
                    // samples [s0, s1) of the slice that the tile's frames read; this warp takes its 1 / NW of them
                    long long s0 = (__ldg(&d2->frame_begin) + fb2) * p.hop - half - p.pad_left - __ldg(&d2->pcm_offset);
                    long long s1 = s0 + (fe2 - fb2 - 1) * p.hop + 2048;
                    s0 = max(s0, 0ll);
                    s1 = min(s1, __ldg(&d2->slice_len));
                    const long long per = ((s1 - s0 + NW - 1) / NW + 31) & ~31ll;
                    const long long a = s0 + per * warp, b = min(a + per, s1);
                    if (b > a) {
                        const unsigned long long base = __ldg(reinterpret_cast<const unsigned long long *>(&d2->pcm));
                        const unsigned long long lo = (base + a * esz) & ~15ull, hi = (base + b * esz) & ~15ull;  // never past the slice
                        if (hi > lo) prefetch_l2_bulk(reinterpret_cast<const void *>(lo), static_cast<unsigned>(hi - lo));
                    }
                }
            }
        }

#endif
        for (long long fa = f_begin + 2 * warp; fa < f_end; fa += 2 * NW) {
            const long long fb = fa + 1;
            cx v[32];
            // ---- load + window: v[n1] = z[32 n1 + lane] of both frames ----
            const long long first_a = (d.frame_begin + fa) * p.hop - half - p.pad_left;  // file index of FFT position 0
            if constexpr (I16 && HS < 0) {
                // 16-bit PCM on an odd hop / odd start (CD audio with the default setting): one 2-byte load per sample
                const short *src_a = reinterpret_cast<const short *>(d.pcm) + (first_a - d.pcm_offset) + 2 * lane;
                const short *src_b = src_a + p.hop;
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                    const float a_re = static_cast<float>(static_cast<int>(__ldg(src_a + 64 * n1)));
                    const float a_im = static_cast<float>(static_cast<int>(__ldg(src_a + 64 * n1 + 1)));
                    const float b_re = static_cast<float>(static_cast<int>(__ldg(src_b + 64 * n1)));
                    const float b_im = static_cast<float>(static_cast<int>(__ldg(src_b + 64 * n1 + 1)));
                    v[n1].re = make_float2(a_re * w.x, b_re * w.x);
                    v[n1].im = make_float2(a_im * w.y, b_im * w.y);
                }
            } else if constexpr (I16) {
                const uint32_t *src_a = reinterpret_cast<const uint32_t *>(reinterpret_cast<const short *>(d.pcm) +
                                                                           (first_a - d.pcm_offset) + 2 * lane);
                const uint32_t *src_b = src_a + (p.hop >> 1);
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    const uint32_t wa = __ldg(src_a + 32 * n1), wb = __ldg(src_b + 32 * n1);
                    const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                    // sign-extend each half (low: 16-bit cast, high: arithmetic shift), I2F converts it (exact)
                    const float a_re = static_cast<float>(static_cast<int>(static_cast<short>(wa & 0xffffu)));
                    const float a_im = static_cast<float>(static_cast<int>(wa) >> 16);
                    const float b_re = static_cast<float>(static_cast<int>(static_cast<short>(wb & 0xffffu)));
                    const float b_im = static_cast<float>(static_cast<int>(wb) >> 16);
                    v[n1].re = make_float2(a_re * w.x, b_re * w.x);
                    v[n1].im = make_float2(a_im * w.y, b_im * w.y);
                }
            } else if constexpr (HS > 0) {
                const float *src_a = d.pcm + (first_a - d.pcm_offset) + 2 * lane;
                float2 r[32 + HS];
#pragma unroll
                for (int n1 = 0; n1 < 32 + HS; n1++) r[n1] = __ldg(reinterpret_cast<const float2 *>(src_a + 64 * n1));
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                    v[n1].re = make_float2(r[n1].x * w.x, r[n1 + HS].x * w.x);
                    v[n1].im = make_float2(r[n1].y * w.y, r[n1 + HS].y * w.y);
                }
            } else if constexpr (HS < 0) {
                const float *src_a = d.pcm + (first_a - d.pcm_offset) + 2 * lane;
                const float *src_b = src_a + p.hop;
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                    v[n1].re = make_float2(__ldg(src_a + 64 * n1) * w.x, __ldg(src_b + 64 * n1) * w.x);
                    v[n1].im = make_float2(__ldg(src_a + 64 * n1 + 1) * w.y, __ldg(src_b + 64 * n1 + 1) * w.y);
                }
            } else {
                const float *src_a = d.pcm + (first_a - d.pcm_offset) + 2 * lane;
                const float *src_b = src_a + p.hop;
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    const float2 xa = __ldg(reinterpret_cast<const float2 *>(src_a + 64 * n1));
                    const float2 xb = __ldg(reinterpret_cast<const float2 *>(src_b + 64 * n1));
                    const float2 w = *reinterpret_cast<const float2 *>(sm.wpad + 64 * n1 + 2 * lane);
                    v[n1].re = make_float2(xa.x * w.x, xb.x * w.x);
                    v[n1].im = make_float2(xa.y * w.y, xb.y * w.y);
                }
            }
            // ---- pass 1: DFT over n1, twiddle, transpose (real plane, then imaginary plane) ----
            dft32p(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; k1++) {
                if (k1) {
                    const float2 w = sm.tw1[(k1 - 1) * 32 + lane];
                    v[perm32(k1)] = cmul_s(v[perm32(k1)], w.x, w.y);
                }
                tile[k1 * kRow + lane] = v[perm32(k1)].re;
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) v[n2].re = tile[lane * kRow + n2];
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 32; k1++) tile[k1 * kRow + lane] = v[perm32(k1)].im;
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) v[n2].im = tile[lane * kRow + n2];
            __syncwarp();
            // ---- pass 2: DFT over n2 -> Z[lane + 32 k2] in v[perm32(k2)] ----
            dft32p(v);
            // ---- real split: 16 (k, 1024 - k) pairs per lane; |X| or dB ----
            float *orow_a = d.out + fa * p.n_bins, *orow_b = d.out + fb * p.n_bins;
            f2 smax = make_float2(0.0f, 0.0f);
            f2 fmx = make_float2(-CUDART_INF_F, -CUDART_INF_F), fnm = fmx;  // this pair's own max / -min (linear)
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const cx own_a = v[perm32(31 - j)], own_b = v[perm32((32 - j) & 31)];
                cx zk, zn;
                zk.re.x = lane ? own_a.re.x : own_b.re.x;
                zk.re.y = lane ? own_a.re.y : own_b.re.y;
                zk.im.x = lane ? own_a.im.x : own_b.im.x;
                zk.im.y = lane ? own_a.im.y : own_b.im.y;
                const cx sup = v[perm32(j)];
                zn.re.x = __shfl_sync(0xffffffffu, sup.re.x, partner);
                zn.re.y = __shfl_sync(0xffffffffu, sup.re.y, partner);
                zn.im.x = __shfl_sync(0xffffffffu, sup.im.x, partner);
                zn.im.y = __shfl_sync(0xffffffffu, sup.im.y, partner);
                const f2 er = padd(zk.re, zn.re), ei = psub(zk.im, zn.im), dr = psub(zk.re, zn.re), di = padd(zk.im, zn.im);
                const float2 w = sm.tw2[j * 32 + lane];
                const f2 wr = pfma(di, bc(-w.y), pmul(dr, bc(w.x))), wi = pfma(dr, bc(w.y), pmul(di, bc(w.x)));
                const f2 ar = padd(er, wi), ai = psub(ei, wr), br = psub(er, wi), bi = padd(ei, wr);
                const f2 sa = pfma(ar, ar, pmul(ai, ai)), sb = pfma(br, br, pmul(bi, bi));
                smax.x = fmaxf(smax.x, fmaxf(sa.x, sb.x));
                smax.y = fmaxf(smax.y, fmaxf(sa.y, sb.y));
                const int k_own = lane32 + 32 * (31 - j), k_par = 1024 - k_own;
                if (MEL) {
                    mag[k_own] = make_float2(sqrt_ftz(sa.x), sqrt_ftz(sa.y));
                    mag[k_par] = make_float2(sqrt_ftz(sb.x), sqrt_ftz(sb.y));
                } else {
                    const f2 a = pmul(make_float2(lg2_ftz(sa.x), lg2_ftz(sa.y)), bc(kDbPerLog2Pow));
                    const f2 b = pmul(make_float2(lg2_ftz(sb.x), lg2_ftz(sb.y)), bc(kDbPerLog2Pow));
                    orow_a[k_own] = a.x;
                    orow_a[k_par] = b.x;
                    orow_b[k_own] = a.y;
                    orow_b[k_par] = b.y;
                    fmx.x = fmaxf(fmx.x, fmaxf(a.x, b.x));
                    fmx.y = fmaxf(fmx.y, fmaxf(a.y, b.y));
                    fnm.x = fmaxf(fnm.x, fmaxf(-a.x, -b.x));
                    fnm.y = fmaxf(fnm.y, fmaxf(-a.y, -b.y));
                }
            }
            if (lane == 0) {  // k = 512 pairs with itself: X[512] = conj(2 Z'[512])
                const cx z = v[perm32(16)];
                const f2 s5 = pmul(pfma(z.re, z.re, pmul(z.im, z.im)), bc(4.0f));
                smax.x = fmaxf(smax.x, s5.x);
                smax.y = fmaxf(smax.y, s5.y);
                if (MEL) {
                    mag[512] = make_float2(sqrt_ftz(s5.x), sqrt_ftz(s5.y));
                } else {
                    const f2 a = pmul(make_float2(lg2_ftz(s5.x), lg2_ftz(s5.y)), bc(kDbPerLog2Pow));
                    orow_a[512] = a.x;
                    orow_b[512] = a.y;
                    fmx.x = fmaxf(fmx.x, a.x);
                    fmx.y = fmaxf(fmx.y, a.y);
                    fnm.x = fmaxf(fnm.x, -a.x);
                    fnm.y = fmaxf(fnm.y, -a.y);
                }
            }
            // frames outside the exact range of the f32 square: the tile goes on the rescue list, and nothing of
            // this pair enters the min/max here (the scalar kernel redoes the whole tile, bit-identically for
            // the frames that were fine)
            const float mx_a = warp_max(smax.x), mx_b = warp_max(smax.y);
            const bool pair_ok = !((mx_a < kPowTiny && mx_a > 0.0f) || mx_a > kPowHuge || (mx_b < kPowTiny && mx_b > 0.0f) ||
                                   mx_b > kPowHuge);
            flagged |= !pair_ok;
            if (!MEL && pair_ok) {
                lmax = fmaxf(lmax, fmaxf(fmx.x, fmx.y));
                lnmin = fmaxf(lnmin, fmaxf(fnm.x, fnm.y));
            }
            if constexpr (PL && HS > 0 && !I16) {
                // the (32 + HS) x 256-byte rows of the warp's next frame pair are contiguous: lane l touches 128-byte lines
                // l, l + 32, l + 64 of them -- three instructions put the whole pair into L1 while the mel stage runs
                const long long fn = fa + 2 * NW;
                if (fn < f_end) {
                    const char *base = reinterpret_cast<const char *>(d.pcm + ((d.frame_begin + fn) * p.hop - half - p.pad_left - d.pcm_offset));
                    constexpr int kLines = (32 + HS) * 2;
#pragma unroll
                    for (int i = 0; i * 32 < kLines; i++)
                        if (i * 32 + lane < kLines) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + 128 * (i * 32 + lane)));
                }
            }
            if (MEL) {
                __syncwarp();
                const MelView mv(sm.ms);
                if constexpr (!DIRECT) {
                    if constexpr (M4) mel_walk4<float2>(mv, mag, part, lane);
                    else mel_walk<float2>(mv, mag, part, lane);
                    __syncwarp();
                }
                // (two rounds per walk -- mel_direct2, as in thb_stft_warp.cu -- were measured here too: 5.04 -> 5.26 ms at the
                // 48 kHz default bank, 5.75 with its padding; this kernel's shared-memory pipe is its busiest unit)
                for (int r = 0; 32 * r < mv.n_mel; r++) {
                    const int m = 32 * r + lane;
                    f2 acc;
                    if constexpr (DIRECT) acc = mel_direct<float2>(mv, mag, r, lane);
                    else acc = M4 ? mel_band4<float2>(mv, part, r, lane) : mel_band<float2>(mv, part, r, lane);
                    if (m >= mv.n_mel) continue;
                    const f2 db = pmul(make_float2(lg2_ftz(acc.x), lg2_ftz(acc.y)), bc(kDbPerLog2Amp));
                    orow_a[m] = db.x;
                    orow_b[m] = db.y;
                    if (pair_ok) {
                        lmax = fmaxf(lmax, fmaxf(db.x, db.y));
                        lnmin = fmaxf(lnmin, fmaxf(-db.x, -db.y));
                    }
                }
                __syncwarp();
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
            if (flagged && atomicExch(&rescue.flags[item], 1u) == 0u) {
                const unsigned slot = atomicAdd(rescue.count, 1u);
                if (slot < rescue.capacity) rescue.items[slot] = make_uint2(static_cast<unsigned>(track), static_cast<unsigned>(tile_idx));
            }
        }
    }
}

size_t pair_smem_bytes(const PlanDev &p, int nw) {
    return sizeof(float) * 2048 + sizeof(float2) * (31 * 32 + 16 * 32) + sizeof(float2) * nw * tile_elems(p) +
           sizeof(uint32_t) * static_cast<size_t>(p.n_mel ? p.mi_words : 0);
}

int pair_warps() {
    // THB_PAIR_WARPS = 8 | 10 | 12 (tuning knob; registers per thread follow from it)
    const char *e = getenv("THB_PAIR_WARPS");
    const int w = e ? atoi(e) : 12;
    return (w == 8 || w == 10 || w == 12 || w == 14 || w == 16) ? w : 12;
}

template <bool MEL, int NW, bool I16, int HS, bool M4, bool PL, bool DIRECT>
cudaError_t launch_nwd(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, RescueList rescue, int sm_count,
                       cudaStream_t st) {
    const size_t smem = pair_smem_bytes(plan, NW);
    auto kern = stft2048_pair_kernel<MEL, NW, I16, HS, M4, PL, DIRECT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const long long n_items = static_cast<long long>(n_tracks) * rescue.tiles_per_track;
    const int grid = static_cast<int>(n_items < sm_count ? n_items : sm_count);
    // THB_PAIR_PREFETCH=1 turns the next-tile L2 prefetch on (A/B runs).  Off by default: measured on B200 it buys
    // nothing on the mel kernel (1.077 -> 1.074 ms at 450 k frames), 1.7 % on the linear one, and ncu shows 19 % more DRAM
    // reads (lines prefetched 30 us ahead are partly evicted again before their tile starts).
    const char *pe = getenv("THB_PAIR_PREFETCH");
    const int flags = (pe && atoi(pe) == 1) ? 1 : 0;
    kern<<<grid, NW * 32, smem, st>>>(plan, d_tracks, n_items, rescue, flags);
    return cudaGetLastError();
}

template <bool MEL, int NW, bool I16, int HS = 0, bool M4 = true, bool PL = false>
cudaError_t launch_nw(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, RescueList rescue, int sm_count,
                      cudaStream_t st) {
    if constexpr (MEL && M4 && !PL) {
        if (plan.mi_direct) return launch_nwd<MEL, NW, I16, HS, true, false, true>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    }
    return launch_nwd<MEL, NW, I16, HS, M4, PL, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
}

}  // namespace

int stft_pair_tile_frames() { return 8 * pair_warps(); }

bool stft_pair_supported(const PlanDev &p) {
    if (p.n_fft != 2048 || !p.fast_wpad || !p.fast_tw) return false;
    if (p.n_mel && (!p.mi_blob || p.mi_min_start < -(kMagBase - 1))) return false;
    return pair_smem_bytes(p, pair_warps()) <= 220 * 1024;
}

// rescue.tile_frames must be stft_pair_tile_frames() and rescue.tiles_per_track = ceil(max n_frames / tile_frames)
cudaError_t launch_stft_pair(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, RescueList rescue,
                             bool pcm_i16, bool unaligned, int sm_count, cudaStream_t st) {
    if (n_tracks <= 0 || rescue.tiles_per_track == 0) return cudaSuccess;
    const int nw = pair_warps();
    if (unaligned && !pcm_i16) {
        if (plan.n_mel) return launch_nw<true, 12, false, -1>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        return launch_nw<false, 12, false, -1>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    }
    if (unaligned) {
        if (plan.n_mel) return launch_nw<true, 12, true, -1>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        return launch_nw<false, 12, true, -1>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    }
    if (pcm_i16) {  // the ingest variant exists for the tuned warp count only
        if (plan.n_mel) return launch_nw<true, 12, true>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        return launch_nw<false, 12, true>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    }
    // THB_PAIR_SHARE=0 turns the shared-load variants off (A/B runs)
    const char *se = getenv("THB_PAIR_SHARE");
    const bool share = !(se && atoi(se) == 0) && nw >= 12;
#ifndef THB_PAIR_EXPERIMENTS
    if (nw > 12) return cudaErrorInvalidValue;  // 14 / 16 warps are only built with -DTHB_PAIR_EXPERIMENTS
#endif
    if (share && plan.hop == 512) {
        // THB_MEL4=0: the mel walk / gather with remainder code and 16-bit slot ids (A/B runs; same results)
        const char *m4 = getenv("THB_MEL4");
        if (plan.n_mel && m4 && atoi(m4) == 0) return launch_nw<true, 12, false, 8, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
#ifdef THB_PAIR_EXPERIMENTS   // the round-2 A/B variants: make NVFLAGS+=-DTHB_PAIR_EXPERIMENTS
        // THB_PAIR_PL=1: the PCM of the warp's next pair prefetched into L1 ahead of the mel stage (3 % slower)
        const char *pl = getenv("THB_PAIR_PL");
        if (plan.n_mel && pl && atoi(pl) == 1) return launch_nw<true, 12, false, 8, true, true>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        // THB_PAIR_WARPS=14 | 16: more warps on fewer registers (128: the DFT state spills; 32 % / 26 % slower)
        if (plan.n_mel && nw == 14) return launch_nw<true, 14, false, 8>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        if (plan.n_mel && nw == 16) return launch_nw<true, 16, false, 8>(plan, d_tracks, n_tracks, rescue, sm_count, st);
#endif
        if (plan.n_mel) return launch_nw<true, 12, false, 8>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        return launch_nw<false, 12, false, 8>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    }
    if (share && plan.hop == 256) {
        if (plan.n_mel) return launch_nw<true, 12, false, 4>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        return launch_nw<false, 12, false, 4>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    }
    if (plan.n_mel) {
        if (nw == 8) return launch_nw<true, 8, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        if (nw == 10) return launch_nw<true, 10, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
        return launch_nw<true, 12, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    }
    if (nw == 8) return launch_nw<false, 8, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    if (nw == 10) return launch_nw<false, 10, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
    return launch_nw<false, 12, false>(plan, d_tracks, n_tracks, rescue, sm_count, st);
}

}  // namespace thb
