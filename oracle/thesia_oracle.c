/* thesia_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the per-track analysis hot path of Sytronik/thesia
 * (SURVEY.md section 8a), written from the algorithm, not copied: plain C99,
 * OpenMP for the batch driver.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product (libthesia_b200.so) never links, loads or calls it.
 *
 * PARITY PIN: every known-answer vector the reference's own tests hold for this
 * path is reproduced by tests/test_oracle_kat.py (stft_works, stft_short_wav,
 * hann_window_works, pad_works, mel_hz_convert, mel_works, mel_default_works,
 * scalar/array dB tests, spectrogram_to_img..., waveform_tile_* and the
 * find_min_max / sum tests).  The reference has NO test pinning real-signal
 * STFT / mel / dB values and cannot be built here (no cargo/rustc), so
 * end-to-end values are pinned only by this restatement plus an independent
 * f64 cross-check against scipy.fft.rfft ("end-to-end parity unpinned by the
 * reference's own tests" -- see DESIGN.md).
 *
 * Two arithmetic variants share one source (orc_fft.inc):
 *   f32  -- "reference-like": every op in f32 in the reference's order
 *           (f32 window, f32 FFT, hypotf, DENSE f32 mel product, log10f, *20).
 *   f64  -- "truth": window / filterbank still generated in f32 exactly like the
 *           reference, then widened; FFT, magnitude, mel, log in double.
 *
 * Reference citations are relative to /root/reference/.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define REAL float
#define SUF f32
#include "orc_fft.inc"
#undef REAL
#undef SUF
#define REAL double
#define SUF f64
#include "orc_fft.inc"
#undef REAL
#undef SUF

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* a1  SpecSetting framing parameters  (src-tauri/src/core/spectrogram.rs:57-98) */
/* ------------------------------------------------------------------------- */
static size_t orc_next_pow2(size_t x) {
    /* usize::next_power_of_two: 0 -> 1 */
    size_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

ORC_API void orc_framing_params(double win_ms, uint32_t sr, uint32_t t_overlap, uint32_t f_overlap,
                                uint64_t *hop, uint64_t *win, uint64_t *n_fft) {
    /* calc_win_length_float: win_ms * sr / 1000 (spectrogram.rs:91-93)
       calc_hop_length: (that / t_overlap).round() as usize (:62-64); Rust round = half away
       from zero = C round(); negative/NaN saturate to 0. */
    double wl = win_ms * (double)sr / 1000.0;
    double h = round(wl / (double)t_overlap);
    uint64_t hh = (h > 0.0 && h == h) ? (h >= 1.8446744073709552e19 ? UINT64_MAX : (uint64_t)h) : 0;
    *hop = hh;
    *win = hh * (uint64_t)t_overlap;                           /* :57-59,85-87 */
    *n_fft = orc_next_pow2((size_t)*win) * (uint64_t)f_overlap; /* :95-97 */
}

/* ------------------------------------------------------------------------- */
/* a2  normalised periodic Hann  (src-tauri/src/core/windows.rs:12-38,68-83)   */
/* ------------------------------------------------------------------------- */
ORC_API void orc_hann_f32(uint64_t size, int symmetric, float *out) {
    /* cosine_window(a=.5, b=.5, c=0, d=0, size, symmetric) in f32 */
    const float a = 0.5f, b = 0.5f, c = 0.0f, d = 0.0f;
    uint64_t size2 = symmetric ? size : size + 1;
    const float pi = (float)M_PI; /* f32::PI */
    for (uint64_t i = 0; i < size; i++) {
        float x = pi * (float)i / (float)(size2 - 1);
        float b_ = b * cosf(2.0f * x);
        float c_ = c * cosf(4.0f * x);
        float d_ = d * cosf(6.0f * x);
        out[i] = (a - b_) + (c_ - d_);
    }
}

ORC_API void orc_normalized_hann_f32(uint64_t win, uint64_t n_fft, float *out) {
    /* calc_normalized_win(Hann, size, norm_factor): hann(size,false) / norm_factor (windows.rs:25) */
    orc_hann_f32(win, 0, out);
    float nf = (float)n_fft;
    for (uint64_t i = 0; i < win; i++) out[i] = out[i] / nf;
}

/* ------------------------------------------------------------------------- */
/* a3  Pad::pad(.., Reflect) along a 1-D array  (src-tauri/src/core/utils.rs:83-141) */
/* ------------------------------------------------------------------------- */
/* i-th element of chain(x.skip(1), x.rev().skip(1)).cycle()   (utils.rs:112-117) */
static float orc_reflect_seq_left(const float *x, size_t n, size_t i) {
    if (n < 2) return x[0]; /* reference: empty cycle -> pad left UNINITIALISED (UB); oracle picks x[0] */
    size_t period = 2 * (n - 1); /* n >= 2 */
    size_t m = i % period;
    /* first n-1 items: x[1..n-1]; next n-1 items: x[n-2], ..., x[0] */
    return (m < n - 1) ? x[1 + m] : x[n - 2 - (m - (n - 1))];
}
/* i-th element of chain(x.rev().skip(1), x.skip(1)).cycle()   (utils.rs:126-131) */
static float orc_reflect_seq_right(const float *x, size_t n, size_t i) {
    if (n < 2) return x[0];
    size_t period = 2 * (n - 1);
    size_t m = i % period;
    return (m < n - 1) ? x[n - 2 - m] : x[1 + (m - (n - 1))];
}

/* out must hold n + pad_left + pad_right. n >= 2 when any pad > 0 (n == 1 is UB in the
   reference: the cycled iterator is empty and the pad stays uninitialised). */
ORC_API void orc_pad_reflect_f32(const float *x, uint64_t n, uint64_t pad_left, uint64_t pad_right,
                                 float *out) {
    memcpy(out + pad_left, x, sizeof(float) * n);
    for (uint64_t i = 0; i < pad_left; i++) out[pad_left - 1 - i] = orc_reflect_seq_left(x, n, i);
    for (uint64_t i = 0; i < pad_right; i++)
        out[pad_left + n + i] = orc_reflect_seq_right(x, n, i);
}

/* ------------------------------------------------------------------------- */
/* a4  perform_stft framing  (src-tauri/src/core/spectrogram/stft.rs:16-149)   */
/* ------------------------------------------------------------------------- */
static uint64_t orc_n_windows(uint64_t len, uint64_t win, uint64_t hop) {
    /* ndarray windows_with_stride: one window per stride position that fits */
    return len >= win ? (len - win) / hop + 1 : 0;
}

/* Frame list as (pointer into a scratch signal, count).  The restatement keeps the
   reference's THREE-PIECE structure (front / mid / back, stft.rs:77-95) so that the
   frame count and edge content are derived the way the reference derives them.
   Output: frames[t*win + i] = raw (un-windowed) samples of frame t.  Returns T. */
ORC_API uint64_t orc_stft_frames_f32(const float *x, uint64_t n, uint64_t win, uint64_t hop,
                                     float *frames /* may be NULL: count only */) {
    uint64_t half = win / 2;
    uint64_t t = 0;
    if (n < win) { /* stft.rs:50-76 */
        uint64_t plen = n + 2 * half;
        float *p = (float *)malloc(sizeof(float) * (plen ? plen : 1));
        orc_pad_reflect_f32(x, n, half, half, p);
        uint64_t nf = orc_n_windows(plen, win, hop);
        if (frames)
            for (uint64_t f = 0; f < nf; f++) memcpy(frames + f * win, p + f * hop, sizeof(float) * win);
        free(p);
        return nf;
    }
    /* front (stft.rs:77-81): input[..win-1] padded (win/2, 0) */
    {
        uint64_t flen = (win - 1) + half;
        float *p = (float *)malloc(sizeof(float) * (flen ? flen : 1));
        orc_pad_reflect_f32(x, win - 1, half, 0, p);
        uint64_t nf = orc_n_windows(flen, win, hop);
        if (frames)
            for (uint64_t f = 0; f < nf; f++) memcpy(frames + (t + f) * win, p + f * hop, sizeof(float) * win);
        t += nf;
        free(p);
    }
    uint64_t first_i = t * hop - half; /* stft.rs:83 */
    /* mid (stft.rs:84-85) */
    {
        uint64_t nm = orc_n_windows(n - first_i, win, hop);
        if (frames)
            for (uint64_t f = 0; f < nm; f++)
                memcpy(frames + (t + f) * win, x + first_i + f * hop, sizeof(float) * win);
        t += nm;
        first_i += nm * hop; /* stft.rs:87 */
    }
    /* back (stft.rs:88-95) */
    {
        uint64_t i_back = first_i < (n - half - 1) ? first_i : (n - half - 1);
        uint64_t blen = (n - i_back) + half;
        float *p = (float *)malloc(sizeof(float) * (blen ? blen : 1));
        orc_pad_reflect_f32(x + i_back, n - i_back, 0, half, p);
        uint64_t skip = first_i > i_back ? first_i - i_back : 0; /* slice_collapse */
        uint64_t nb = blen >= skip ? orc_n_windows(blen - skip, win, hop) : 0;
        if (frames)
            for (uint64_t f = 0; f < nb; f++)
                memcpy(frames + (t + f) * win, p + skip + f * hop, sizeof(float) * win);
        t += nb;
        free(p);
    }
    return t;
}

/* Closed form the product uses (checked against the 3-piece version in tests):
   T = 1 + floor((N + 2*floor(W/2) - W) / H), and frame t, tap i reads
   reflect(t*H + i - floor(W/2)). */
ORC_API uint64_t orc_n_frames(uint64_t n, uint64_t win, uint64_t hop) {
    uint64_t plen = n + 2 * (win / 2);
    return plen >= win ? (plen - win) / hop + 1 : 0;
}

ORC_API int64_t orc_reflect_index(int64_t s, int64_t n) {
    /* numpy-style 'reflect' with period 2n-2 (n >= 2) */
    int64_t period = 2 * (n - 1);
    int64_t m = s % period;
    if (m < 0) m += period;
    return m < n ? m : period - m;
}

/* ------------------------------------------------------------------------- */
/* a7  mel scale + filterbank  (src-common/src/lib.rs:11-103)                  */
/* ------------------------------------------------------------------------- */
#define ORC_MIN_LOG_MEL 15
static const double ORC_MIN_LOG_HZ = 1000.;
static const double ORC_LOGSTEP = 0.06875177742094912;
static const double ORC_LINEARSCALE = 200. / 3.;

ORC_API float orc_mel_to_hz_f32(float mel) { /* lib.rs:18-29 */
    float min_log_mel = (float)ORC_MIN_LOG_MEL;
    if (mel < min_log_mel) return (float)ORC_LINEARSCALE * mel;
    return (float)ORC_MIN_LOG_HZ * expf((float)ORC_LOGSTEP * (mel - min_log_mel));
}
ORC_API float orc_mel_from_hz_f32(float hz) { /* lib.rs:32-43 */
    float min_log_hz = (float)ORC_MIN_LOG_HZ;
    if (hz < min_log_hz) return hz / (float)ORC_LINEARSCALE;
    return (float)ORC_MIN_LOG_MEL + logf(hz / min_log_hz) / (float)ORC_LOGSTEP;
}
ORC_API double orc_mel_to_hz_f64(double mel) {
    if (mel < (double)ORC_MIN_LOG_MEL) return ORC_LINEARSCALE * mel;
    return ORC_MIN_LOG_HZ * exp(ORC_LOGSTEP * (mel - (double)ORC_MIN_LOG_MEL));
}
ORC_API double orc_mel_from_hz_f64(double hz) {
    if (hz < ORC_MIN_LOG_HZ) return hz / ORC_LINEARSCALE;
    return (double)ORC_MIN_LOG_MEL + log(hz / ORC_MIN_LOG_HZ) / ORC_LOGSTEP;
}

/* ndarray numeric_util::unrolled_fold restated (8 partial sums) -- used by Array::sum on a
   contiguous row (lib.rs:85). */
#define ORC_DEF_USUM(T, NAME)                                              \
    static T NAME(const T *xs, size_t n) {                                 \
        T acc = 0, p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0; \
        while (n >= 8) {                                                   \
            p0 += xs[0]; p1 += xs[1]; p2 += xs[2]; p3 += xs[3];            \
            p4 += xs[4]; p5 += xs[5]; p6 += xs[6]; p7 += xs[7];            \
            xs += 8; n -= 8;                                               \
        }                                                                  \
        acc = acc + (p0 + p4); acc = acc + (p1 + p5);                      \
        acc = acc + (p2 + p6); acc = acc + (p3 + p7);                      \
        for (size_t i = 0; i < n && i < 7; i++) acc = acc + xs[i];         \
        return acc;                                                        \
    }
ORC_DEF_USUM(float, orc_usum_f32)
ORC_DEF_USUM(double, orc_usum_f64)

/* calc_mel_fb (lib.rs:46-89): out is (n_fft/2+1, n_mel) row-major.
   linspace(a,b,n)[i] = a + ((b-a)/(n-1)) * i  (ndarray Linspace). */
#define ORC_DEF_MELFB(T, NAME, FROMHZ, TOHZ, USUM, EPS)                                      \
    ORC_API void NAME(uint32_t sr, uint64_t n_fft, uint64_t n_mel, T fmin, int has_fmax,     \
                      T fmax_in, int do_norm, T *out) {                                      \
        T f_nyquist = (T)((double)sr / 2.);                                                  \
        T fmax = has_fmax ? fmax_in : f_nyquist;                                             \
        uint64_t n_freq = n_fft / 2 + 1;                                                     \
        T *lin = (T *)malloc(sizeof(T) * n_freq);                                            \
        T *melf = (T *)malloc(sizeof(T) * (n_mel + 2));                                      \
        T *w = (T *)malloc(sizeof(T) * n_freq);                                              \
        {                                                                                    \
            T step = n_freq > 1 ? (f_nyquist - (T)0) / (T)(n_freq - 1) : (T)0;               \
            for (uint64_t i = 0; i < n_freq; i++) lin[i] = (T)0 + step * (T)i;               \
            T m0 = FROMHZ(fmin), m1 = FROMHZ(fmax);                                          \
            T mstep = (n_mel + 2) > 1 ? (m1 - m0) / (T)(n_mel + 1) : (T)0;                   \
            for (uint64_t i = 0; i < n_mel + 2; i++) melf[i] = TOHZ(m0 + mstep * (T)i);      \
        }                                                                                    \
        for (uint64_t im = 0; im < n_mel; im++) {                                            \
            memset(w, 0, sizeof(T) * n_freq);                                                \
            for (uint64_t i_f = 0; i_f < n_freq; i_f++) {                                    \
                T f = lin[i_f];                                                              \
                if (f <= melf[im]) continue;                                                 \
                else if (melf[im] < f && f < melf[im + 1])                                   \
                    w[i_f] = (f - melf[im]) / (melf[im + 1] - melf[im]);                     \
                else if (f == melf[im + 1]) w[i_f] = (T)1;                                   \
                else if (melf[im + 1] < f && f < melf[im + 2])                               \
                    w[i_f] = (melf[im + 2] - f) / (melf[im + 2] - melf[im + 1]);             \
                else break;                                                                  \
            }                                                                                \
            if (do_norm) {                                                                   \
                T s = USUM(w, n_freq);                                                       \
                if (!(s > EPS)) s = EPS; /* .max(A::epsilon()) */                            \
                for (uint64_t i_f = 0; i_f < n_freq; i_f++) w[i_f] /= s;                     \
            }                                                                                \
            for (uint64_t i_f = 0; i_f < n_freq; i_f++) out[i_f * n_mel + im] = w[i_f];      \
        }                                                                                    \
        free(lin); free(melf); free(w);                                                      \
    }
ORC_DEF_MELFB(float, orc_mel_fb_f32, orc_mel_from_hz_f32, orc_mel_to_hz_f32, orc_usum_f32, FLT_EPSILON)
ORC_DEF_MELFB(double, orc_mel_fb_f64, orc_mel_from_hz_f64, orc_mel_to_hz_f64, orc_usum_f64, DBL_EPSILON)

static int orc_melfb_all_nonempty_f32(const float *fb, uint64_t n_freq, uint64_t n_mel) {
    /* mel_fb.sum_axis(Axis(0)).iter().all(|&x| x > 0.)  (lib.rs:98) */
    for (uint64_t m = 0; m < n_mel; m++) {
        float s = 0.f;
        for (uint64_t k = 0; k < n_freq; k++) s += fb[k * n_mel + m];
        if (!(s > 0.f)) return 0;
    }
    return 1;
}

/* calc_mel_fb_default (lib.rs:91-103): returns n_mel; if out != NULL it must hold
   (n_fft/2+1) * min(initial n_mel, n_fft/2+1) floats and receives the (n_freq, n_mel) bank. */
ORC_API uint64_t orc_mel_fb_default_f32(uint32_t sr, uint64_t n_fft, float *out) {
    float v = fmaf(orc_mel_from_hz_f32((float)sr / 2.f) / orc_mel_from_hz_f32((float)sr / (float)n_fft),
                   2.f, -1.f);
    uint64_t n_mel = (v > 0.f && v == v) ? (uint64_t)v : 0; /* `as usize`: trunc, saturating */
    uint64_t n_freq = n_fft / 2 + 1;
    if (n_mel > n_freq) n_mel = n_freq;
    float *fb = (float *)malloc(sizeof(float) * n_freq * (n_mel ? n_mel : 1));
    for (;;) {
        orc_mel_fb_f32(sr, n_fft, n_mel, 0.f, 0, 0.f, 1, fb);
        if (orc_melfb_all_nonempty_f32(fb, n_freq, n_mel)) break;
        n_mel -= 1;
    }
    if (out) memcpy(out, fb, sizeof(float) * n_freq * n_mel);
    free(fb);
    return n_mel;
}

/* a13 FreqScale::hz_range_to_idx (lib.rs:135-159); freq_scale: 0 Linear, 1 Mel */
ORC_API void orc_hz_range_to_idx(int freq_scale, float hz0, float hz1, uint32_t sr, uint64_t n_bins,
                                 uint64_t *i0, uint64_t *i1) {
    if (hz0 >= hz1) { *i0 = 0; *i1 = 0; return; }
    float half_sr = (float)sr / 2.f;
    float r0, r1;
    if (freq_scale == 0) { r0 = hz0 / half_sr; r1 = hz1 / half_sr; }
    else {
        r0 = orc_mel_from_hz_f32(hz0) / orc_mel_from_hz_f32(half_sr);
        r1 = orc_mel_from_hz_f32(hz1) / orc_mel_from_hz_f32(half_sr);
    }
    float a = floorf(r0 * (float)n_bins);
    if (!(a > 0.f)) a = 0.f; /* .max(0.) */
    float b = ceilf(r1 * (float)n_bins);
    *i0 = (uint64_t)a;
    *i1 = (b > 0.f) ? (uint64_t)b : 0;
}

/* ------------------------------------------------------------------------- */
/* a9  dB  (src-tauri/src/core/dynamics/decibel.rs:170-214)                    */
/* ------------------------------------------------------------------------- */
/* generic scalar rule (decibel.rs:66-89 / 176-195), factor applied afterwards (:198-202) */
ORC_API float orc_dB_scalar_f32(float x, float ref_value, float amin, float factor) {
    if (ref_value != ref_value || signbit(ref_value)) return NAN;
    float log_amin = log10f(amin);
    float log_ref = ref_value > amin ? log10f(ref_value) : log_amin;
    float out_for_small = log_amin - log_ref;
    float y;
    if (x != x || signbit(x)) y = NAN;
    else if (x > amin) y = log10f(x) - log_ref;
    else y = out_for_small;
    return y * factor;
}
static inline double orc_dB_amp_default_f64(double x) {
    if (x != x || signbit(x)) return NAN;
    if (x > 0.0) return log10(x) * 20.0;
    return -INFINITY;
}
ORC_API void orc_dB_from_amp_inplace_f32(float *x, uint64_t n, float ref_value, float amin) {
    /* two passes like the reference: log_for_dB_inplace then scalar_mul(20) */
    for (uint64_t i = 0; i < n; i++) x[i] = orc_dB_scalar_f32(x[i], ref_value, amin, 1.0f);
    for (uint64_t i = 0; i < n; i++) x[i] *= 20.0f;
}

/* ------------------------------------------------------------------------- */
/* a12/a17  find_min_max, sum  (src-tauri/src/core/simd.rs:14-36,137-159)      */
/* ------------------------------------------------------------------------- */
ORC_API void orc_find_min_max_f32(const float *x, uint64_t n, float *mn, float *mx) {
    /* f32::min / f32::max ignore NaN; empty -> (+inf, -inf) (simd.rs:274-281) */
    float a = INFINITY, b = -INFINITY;
    for (uint64_t i = 0; i < n; i++) { a = fminf(a, x[i]); b = fmaxf(b, x[i]); }
    *mn = a; *mx = b;
}

/* sum_squares_scalar (simd.rs:820-832): Kahan-compensated sum of x*x in f32.  The reference's SIMD legs keep one
   compensated accumulator per lane (alignment dependent), so only the scalar leg is restated; tests compare within
   f32 rounding. */
ORC_API float orc_sum_squares_f32(const float *x, uint64_t n) {
    volatile float sum = 0.0f, c = 0.0f;  /* volatile: keeps the compiler from simplifying the compensation away */
    for (uint64_t i = 0; i < n; i++) {
        volatile float y = x[i] * x[i] - c;
        volatile float t = sum + y;
        c = (t - sum) - y;
        sum = t;
    }
    return sum;
}

/* abs_max_scalar (simd.rs:935-937): fold(0, max) over |x| (empty -> 0) */
ORC_API float orc_abs_max_f32(const float *x, uint64_t n) {
    float m = 0.0f;
    for (uint64_t i = 0; i < n; i++) {
        const float a = fabsf(x[i]);
        if (a > m) m = a;
    }
    return m;
}

/* StatCalculator::calc without the loudness leg (dynamics/stats.rs:56-85): wavs is (n_ch, n) row-major.
   mean_squared = sum over channels of sum_squares(channel) / n_elem (f32), rms_dB = 10 log10 (dB_from_power_default,
   decibel.rs:95-107: 0 -> -inf), max_peak = abs_max over everything, max_peak_dB = 20 log10.  out = {mean_squared,
   rms_dB, max_peak, max_peak_dB}. */
ORC_API void orc_audio_stats_f32(const float *wavs, uint64_t n_ch, uint64_t n, float out[4]) {
    float total = 0.0f, peak = 0.0f;
    for (uint64_t c = 0; c < n_ch; c++) {
        total += orc_sum_squares_f32(wavs + c * n, n);
        const float m = orc_abs_max_f32(wavs + c * n, n);
        if (m > peak) peak = m;
    }
    const float ms = total / (float)(n_ch * n);
    out[0] = ms;
    out[1] = 10.0f * log10f(ms);
    out[2] = peak;
    out[3] = 20.0f * log10f(peak);
}

/* ------------------------------------------------------------------------- */
/* f4  gain normalisation + guard clipping           */
/* ------------------------------------------------------------------------- */
/* Normalize::normalize_default (dynamics/normalize.rs:23-45): the gain of a target, from the ORIGINAL audio's
   stats.  kind 0 Off, 1 LUFS, 2 RMSdB, 3 PeakdB.  `global_lufs` is f64 in the reference and narrowed first. */
ORC_API float orc_normalize_gain(int kind, float target, double global_lufs, float rms_dB, float max_peak_dB) {
    switch (kind) {
    case 1: return powf(10.0f, (target - (float)global_lufs) / 20.0f);
    case 2: return powf(10.0f, (target - rms_dB) / 20.0f);
    case 3: return powf(10.0f, (target - max_peak_dB) / 20.0f);
    default: return 1.0f;
    }
}

static inline float orc_clamp_unit(float x) { /* f32::clamp(-1, 1): NaN stays NaN */
    if (x < -1.0f) return -1.0f;
    if (x > 1.0f) return 1.0f;
    return x;
}

/* AudioTrack::apply_gain (track.rs:152-171) followed by Audio::mutate's guard clipping (audio.rs:49-63) for the
   two elementwise modes: 0 Clip (audio.rs:134-144), 1 ReduceGlobalLevel (audio.rs:146-160).  wavs is (n_ch, n)
   row-major = AudioTrack.original; out receives AudioTrack.audio.wavs; before_clip (may be NULL) receives
   GuardClippingResult::WavBeforeClip in mode 0.  *global_gain = the GlobalGain of mode 1 (1 otherwise).
   gc_dB / gc_cnt [n_ch] = GuardClippingStats {max_reduction_gain_dB, reduction_cnt} per channel
   (dynamics/stats.rs:133-158,176-186).  A non-finite or unit gain restores the original (track.rs:160-161),
   whose guard-clip state is GlobalGain(1) (audio.rs:33-44).  Returns 0, or -1 for the limiter mode (a sequential
   recurrence, not restated). */
ORC_API int orc_apply_gain(const float *wavs, uint64_t n_ch, uint64_t n, float gain, int mode, float *out,
                           float *before_clip, float *global_gain, float *gc_dB, uint64_t *gc_cnt) {
    *global_gain = 1.0f;
    if (!isfinite(gain) || gain == 1.0f) {
        memcpy(out, wavs, sizeof(float) * n_ch * n);
        for (uint64_t c = 0; c < n_ch; c++) { gc_dB[c] = log10f(1.0f) * 20.0f; gc_cnt[c] = 0; }
        return 0;
    }
    if (mode != 0 && mode != 1) return -1;
    for (uint64_t i = 0; i < n_ch * n; i++) out[i] = gain * wavs[i];     /* azip!(*y = gain * x) */
    if (mode == 0) {
        if (before_clip) memcpy(before_clip, out, sizeof(float) * n_ch * n);
        for (uint64_t c = 0; c < n_ch; c++) {
            float *y = out + c * n;
            const float peak = orc_abs_max_f32(y, n);                   /* from_wav_before_clip */
            gc_dB[c] = 0.0f; gc_cnt[c] = 0;
            if (peak > 1.0f) {
                gc_dB[c] = log10f(1.0f / peak) * 20.0f;
                for (uint64_t i = 0; i < n; i++) gc_cnt[c] += fabsf(y[i]) > 1.0f;
            }
            for (uint64_t i = 0; i < n; i++) y[i] = orc_clamp_unit(y[i]);
        }
    } else {
        const double peak = (double)orc_abs_max_f32(out, n_ch * n);      /* wavs.max_peak() as f64 */
        float g32 = 1.0f;
        if (peak > 1.0) {
            const double g = 1.0 / peak;
            for (uint64_t i = 0; i < n_ch * n; i++) out[i] = orc_clamp_unit((float)((double)out[i] * g));
            g32 = (float)g;
        }
        *global_gain = g32;
        for (uint64_t c = 0; c < n_ch; c++) { gc_dB[c] = log10f(g32) * 20.0f; gc_cnt[c] = 0; }  /* from_global_gain */
    }
    return 0;
}

/* sum_avx2 order (simd.rs:594-619): scalar prefix up to 32-byte alignment, 8 lane
   accumulators over the aligned middle, lanes reduced, scalar suffix. `align_elems` = number of
   prefix elements (0..7) -- the reference derives it from the slice ADDRESS, so the sum is
   alignment-dependent; tests sweep it. */
ORC_API float orc_sum_simd_order_f32(const float *x, uint64_t n, uint32_t align_elems) {
    if (n == 0) return 0.f;
    float sum = 0.f;
    uint64_t pre = align_elems < n ? align_elems : n;
    uint64_t mid = (n - pre) / 8;
    if (mid == 0) { /* align_to may return everything as prefix */
        for (uint64_t i = 0; i < n; i++) sum += x[i];
        return sum;
    }
    for (uint64_t i = 0; i < pre; i++) sum += x[i];
    float l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float *m = x + pre;
    for (uint64_t c = 0; c < mid; c++)
        for (int j = 0; j < 8; j++) l[j] += m[c * 8 + j];
    /* _mm256_reduce_add_ps (simd.rs:812-817): lo128 + hi128, then _mm_reduce_add_ps
       (simd.rs:800-806): movehdup/add, movehl/add  =>  (q0+q1) + (q2+q3) */
    float q0 = l[0] + l[4], q1 = l[1] + l[5], q2 = l[2] + l[6], q3 = l[3] + l[7];
    sum += (q0 + q1) + (q2 + q3);
    for (uint64_t i = pre + mid * 8; i < n; i++) sum += x[i];
    return sum;
}

/* ------------------------------------------------------------------------- */
/* a16 encode_waveform_tile  (src-tauri/src/core/render_tiles.rs:14,18,232-279) */
/* ------------------------------------------------------------------------- */
#define ORC_TILE_BINS 1024u
static void orc_put_u32(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static void orc_put_f32(uint8_t *p, float f) { uint32_t v; memcpy(&v, &f, 4); orc_put_u32(p, v); }

/* returns number of bytes written (24 + 12*bin_count); out may be NULL to query the size */
ORC_API uint64_t orc_encode_waveform_tile(const float *wav, uint64_t len, uint64_t revision,
                                          uint32_t level, uint32_t tile_index, uint8_t *out) {
    uint64_t spb = level < 64 ? ((uint64_t)1 << level) : UINT64_MAX; /* checked_shl */
    uint64_t tile_samples = (spb > UINT64_MAX / ORC_TILE_BINS) ? UINT64_MAX : ORC_TILE_BINS * spb;
    uint64_t start = (tile_samples != 0 && (uint64_t)tile_index > UINT64_MAX / tile_samples)
                         ? UINT64_MAX : (uint64_t)tile_index * tile_samples;
    uint64_t end_ = start > UINT64_MAX - tile_samples ? UINT64_MAX : start + tile_samples;
    uint64_t end = len < end_ ? len : end_;
    uint64_t bin_count = start >= end ? 0 : ((end - start) + spb - 1) / spb;
    if (!out) return 24 + 12 * bin_count;
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)(revision >> (8 * i));
    orc_put_u32(out + 8, (uint32_t)bin_count);
    orc_put_u32(out + 12, (uint32_t)(spb < 0xFFFFFFFFull ? spb : 0xFFFFFFFFull));
    orc_put_u32(out + 16, tile_index);
    orc_put_u32(out + 20, 0);
    for (uint64_t b = 0; b < bin_count; b++) {
        uint64_t bs = start + b * spb;
        uint64_t be = bs > UINT64_MAX - spb ? UINT64_MAX : bs + spb;
        if (be > end) be = end;
        const float *s = wav + bs;
        uint64_t n = be - bs;
        float mn, mx, sum;
        if (n >= 32) { /* WAVEFORM_SIMD_MIN_MAX_THRESHOLD: simd find_min_max + simd sum */
            orc_find_min_max_f32(s, n, &mn, &mx);
            uint32_t mis = (uint32_t)(((uintptr_t)s % 32u) / 4u);
            uint32_t pre = mis ? 8u - mis : 0u;
            sum = orc_sum_simd_order_f32(s, n, pre);
        } else {
            mn = INFINITY; mx = -INFINITY; sum = 0.f;
            for (uint64_t i = 0; i < n; i++) { mn = fminf(mn, s[i]); mx = fmaxf(mx, s[i]); sum += s[i]; }
        }
        float rep = sum / (float)n;
        orc_put_f32(out + 24 + 12 * b, mn);
        orc_put_f32(out + 28 + 12 * b, mx);
        orc_put_f32(out + 32 + 12 * b, rep);
    }
    return 24 + 12 * bin_count;
}

/* ------------------------------------------------------------------------- */
/* f2  encode_spectrogram_tile  (src-tauri/src/core/render_tiles.rs:15-16,281-393) */
/* ------------------------------------------------------------------------- */
/* The resize itself is THIRD-PARTY arithmetic that is not in the checkout: fast_image_resize 6.0.0
   (src-tauri/Cargo.toml:32), Resizer::resize_typed with ResizeAlg::Convolution(FilterType::Lanczos3) on
   pixels::U16 and a fractional crop box (render_tiles.rs:352-393).  PARITY UNPINNED for the resampled values: this
   is a restatement of that crate's published algorithm (the Pillow-SIMD scheme it documents): per axis, f64
   coefficients of the filter stretched by max(scale, 1), bounds trimmed of leading / trailing zero weights,
   normalised to sum 1, quantised to i32 fixed point with the largest precision that keeps the largest weight inside
   i32, i64 accumulation seeded with half an LSB, arithmetic shift, clamp to u16; horizontal pass over the rows the
   vertical pass needs, then the vertical pass.  Cross-checked against an independent implementation of the same scheme:
   Pillow's Lanczos resize of the same crop box in float32 agrees within rounding (<= 0.5 LSB with one resampled
   axis, <= 1.05 LSB with two: the U16 path rounds the intermediate image, Pillow's float path does not) --
   tests/test_oracle_kat.py::test_spectrogram_tile_resampler_against_pillow.  The reference's own tests pin the layout, the LOD / gutter
   arithmetic, the row flip and saturated values (render_tiles.rs:435-471), all reproduced by
   tests/test_oracle_kat.py. */
#define ORC_SPEC_TILE 512u
#define ORC_SPEC_GUTTER 4u

static double orc_sinc(double x) {
    if (x == 0.0) return 1.0;
    x *= M_PI;
    return sin(x) / x;
}
static double orc_lanczos3(double x) { return (x >= -3.0 && x < 3.0) ? orc_sinc(x) * orc_sinc(x / 3.0) : 0.0; }

/* one axis: bounds start[out], size[out], i32 weights w[out * window]; returns window, *precision */
typedef struct { uint32_t *start, *size; int32_t *w; uint32_t window; uint32_t precision; uint32_t n; } orc_axis;

static void orc_axis_free(orc_axis *a) { free(a->start); free(a->size); free(a->w); memset(a, 0, sizeof *a); }

static void orc_axis_coeffs(uint32_t in_size, double in0, double in1, uint32_t out_size, orc_axis *a) {
    memset(a, 0, sizeof *a);
    a->n = out_size;
    const double scale = (in1 - in0) / (double)out_size;
    const double filter_scale = scale > 1.0 ? scale : 1.0;
    const double radius = 3.0 * filter_scale;
    const uint32_t window = (uint32_t)ceil(radius) * 2u + 1u;
    const double recip = 1.0 / filter_scale;
    double *c = (double *)calloc((size_t)window * out_size, sizeof(double));
    a->start = (uint32_t *)calloc(out_size, sizeof(uint32_t));
    a->size = (uint32_t *)calloc(out_size, sizeof(uint32_t));
    a->w = (int32_t *)calloc((size_t)window * out_size, sizeof(int32_t));
    a->window = window;
    double max_w = 0.0;
    for (uint32_t o = 0; o < out_size; o++) {
        const double in_center = in0 + ((double)o + 0.5) * scale;
        double lo = floor(in_center - radius), hi = ceil(in_center + radius);
        if (lo < 0.0) lo = 0.0;
        if (hi > (double)in_size) hi = (double)in_size;
        const uint32_t x_min = (uint32_t)lo, x_max = (uint32_t)hi;
        const double center = in_center - 0.5;
        double *row = c + (size_t)o * window;
        uint32_t n = 0, b0 = x_min, b1 = x_max;
        double ww = 0.0;
        for (uint32_t x = x_min; x < x_max; x++) {
            const double w = orc_lanczos3(((double)x - center) * recip);
            if (x == b0 && w == 0.0) b0++;          /* no zero weights at the start of the bound */
            else { row[n++] = w; ww += w; }
        }
        for (uint32_t i = n; i-- > 0;) {            /* ... nor at its end */
            if (b1 <= b0 || row[i] != 0.0) break;
            b1--;
        }
        if (ww != 0.0) for (uint32_t i = 0; i < n; i++) row[i] /= ww;
        a->start[o] = b0;
        a->size[o] = b1 - b0;
        for (uint32_t i = 0; i < window; i++) if (row[i] > max_w) max_w = row[i];
    }
    /* Normalizer32: precision = the last p for which round(max_w * 2^(p+1)) still fits below 2^31 */
    uint32_t precision = 0;
    for (uint32_t p = 0; p < 31; p++) {
        precision = p;
        const double next = round(max_w * (double)((int64_t)1 << (p + 1)));
        if (next >= 2147483648.0) break;
    }
    a->precision = precision;
    const double q = (double)((int64_t)1 << precision);
    for (size_t i = 0; i < (size_t)window * out_size; i++) a->w[i] = (int32_t)round(c[i] * q);
    free(c);
}

static inline uint16_t orc_fixed_clip(int64_t ss, uint32_t precision) {
    const int64_t v = ss >> precision;
    return (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
}

/* resize_spectrogram_tile (render_tiles.rs:352-393): the (width, height) window at (start_x, start_y) of the
   (lod_height, lod_width) level of detail of img (H, W) row-major u16 */
ORC_API void orc_resize_spectrogram_tile(const uint16_t *img, uint64_t H, uint64_t W, uint64_t lod_w, uint64_t lod_h,
                                         uint64_t start_x, uint64_t start_y, uint64_t width, uint64_t height,
                                         uint16_t *out) {
    const double left = (double)start_x * (double)W / (double)lod_w, top = (double)start_y * (double)H / (double)lod_h;
    const double right = (double)(start_x + width) * (double)W / (double)lod_w;
    const double bottom = (double)(start_y + height) * (double)H / (double)lod_h;
    /* options.crop(left, top, right - left, bottom - top): the box handed to the resizer is (left, top, w, h) */
    const double cw = right - left, chh = bottom - top;
    orc_axis ax, ay;
    orc_axis_coeffs((uint32_t)W, left, left + cw, (uint32_t)width, &ax);
    orc_axis_coeffs((uint32_t)H, top, top + chh, (uint32_t)height, &ay);
    const uint32_t y_first = ay.start[0];
    const uint32_t y_last = ay.start[height - 1] + ay.size[height - 1];
    const uint32_t tmp_h = y_last - y_first;
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)tmp_h * width);
    for (uint32_t y = 0; y < tmp_h; y++) {
        const uint16_t *row = img + (size_t)(y_first + y) * W;
        for (uint32_t x = 0; x < width; x++) {
            int64_t ss = (int64_t)1 << (ax.precision - 1);
            const int32_t *k = ax.w + (size_t)x * ax.window;
            for (uint32_t i = 0; i < ax.size[x]; i++) ss += (int64_t)row[ax.start[x] + i] * (int64_t)k[i];
            tmp[(size_t)y * width + x] = orc_fixed_clip(ss, ax.precision);
        }
    }
    for (uint32_t y = 0; y < height; y++) {
        const int32_t *k = ay.w + (size_t)y * ay.window;
        const uint32_t y0 = ay.start[y] - y_first;
        for (uint32_t x = 0; x < width; x++) {
            int64_t ss = (int64_t)1 << (ay.precision - 1);
            for (uint32_t i = 0; i < ay.size[y]; i++) ss += (int64_t)tmp[(size_t)(y0 + i) * width + x] * (int64_t)k[i];
            out[(size_t)y * width + x] = orc_fixed_clip(ss, ay.precision);
        }
    }
    free(tmp);
    orc_axis_free(&ax);
    orc_axis_free(&ay);
}

static uint64_t orc_sat_mul(uint64_t a, uint64_t b) { return (b && a > UINT64_MAX / b) ? UINT64_MAX : a * b; }
static uint64_t orc_sat_sub(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }
static uint64_t orc_min_u64(uint64_t a, uint64_t b) { return a < b ? a : b; }

/* geometry of a tile (render_tiles.rs:290-312): geo = {lod_width, lod_height, origin_x, origin_y, width, height} */
ORC_API void orc_spectrogram_tile_geometry(uint64_t H, uint64_t W, uint32_t level_x, uint32_t level_y, uint32_t tile_x,
                                           uint32_t tile_y, uint64_t geo[6]) {
    const uint64_t sx = level_x < 64 ? ((uint64_t)1 << level_x) : UINT64_MAX;   /* checked_shl */
    const uint64_t sy = level_y < 64 ? ((uint64_t)1 << level_y) : UINT64_MAX;
    const uint64_t lod_w = W / sx + (W % sx != 0), lod_h = H / sy + (H % sy != 0);
    const uint64_t start_x = orc_sat_mul(tile_x, ORC_SPEC_TILE), start_y = orc_sat_mul(tile_y, ORC_SPEC_TILE);
    const uint64_t core_w = orc_min_u64(orc_sat_sub(lod_w, start_x), ORC_SPEC_TILE);
    const uint64_t core_h = orc_min_u64(orc_sat_sub(lod_h, start_y), ORC_SPEC_TILE);
    const uint64_t ox = orc_sat_sub(start_x, ORC_SPEC_GUTTER), oy = orc_sat_sub(start_y, ORC_SPEC_GUTTER);
    uint64_t w = 0, h = 0;
    if (core_w && core_h) {
        w = orc_sat_sub(orc_min_u64(lod_w, start_x + core_w + ORC_SPEC_GUTTER), ox);
        h = orc_sat_sub(orc_min_u64(lod_h, start_y + core_h + ORC_SPEC_GUTTER), oy);
    }
    geo[0] = lod_w; geo[1] = lod_h; geo[2] = ox; geo[3] = oy; geo[4] = w; geo[5] = h;
}

/* encode_spectrogram_tile (render_tiles.rs:281-350): 40-byte header {u64 revision, u32 width, height, level_x,
   level_y, tile_x, tile_y, origin_x, origin_y} then RGBA rows, LAST row of the tile first (high frequencies first).
   Returns the byte count; out == NULL queries it. */
ORC_API uint64_t orc_encode_spectrogram_tile(const uint16_t *img, uint64_t H, uint64_t W, const uint8_t *colormap_rgba,
                                             uint64_t colormap_bytes, uint64_t revision, uint32_t level_x,
                                             uint32_t level_y, uint32_t tile_x, uint32_t tile_y, uint8_t *out) {
    uint64_t g[6];
    orc_spectrogram_tile_geometry(H, W, level_x, level_y, tile_x, tile_y, g);
    const uint64_t width = g[4], height = g[5];
    if (!out) return 40 + width * height * 4;
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)(revision >> (8 * i));
    orc_put_u32(out + 8, (uint32_t)width);
    orc_put_u32(out + 12, (uint32_t)height);
    orc_put_u32(out + 16, level_x);
    orc_put_u32(out + 20, level_y);
    orc_put_u32(out + 24, tile_x);
    orc_put_u32(out + 28, tile_y);
    orc_put_u32(out + 32, (uint32_t)g[2]);
    orc_put_u32(out + 36, (uint32_t)g[3]);
    if (!width || !height) return 40;
    uint16_t *px = (uint16_t *)malloc(sizeof(uint16_t) * width * height);
    orc_resize_spectrogram_tile(img, H, W, g[0], g[1], g[2], g[3], width, height, px);
    const uint64_t colors = colormap_bytes / 4;
    uint8_t *o = out + 40;
    for (uint64_t r = height; r-- > 0;) {
        for (uint64_t x = 0; x < width; x++) {
            const uint64_t v = px[r * width + x];
            const uint64_t ci = colors <= 1 ? 0 : (v * (colors - 1) + 65535u / 2u) / 65535u;
            memcpy(o, colormap_rgba + ci * 4, 4);
            o += 4;
        }
    }
    free(px);
    return 40 + width * height * 4;
}

/* ------------------------------------------------------------------------- */
/* a14 convert_spectrogram_to_img  (src-tauri/src/core/visualize/drawing.rs:4-33) */
/* ------------------------------------------------------------------------- */
/* spec: (T, B) row-major dB.  out: (i1 - i0, T) row-major u16. */
ORC_API void orc_spec_to_img(const float *spec, uint64_t T, uint64_t B, uint64_t i0, uint64_t i1,
                             float dB_min, float dB_max, int has_cmap_len, uint32_t colormap_length,
                             uint16_t *out) {
    uint64_t height = i1 - i0, width = T;
    float dB_span = dB_max - dB_min;
    if (dB_min == dB_max && dB_max == -INFINITY) {
        memset(out, 0, sizeof(uint16_t) * height * width);
        return;
    }
    uint16_t min_value = 1;
    if (has_cmap_len) {
        double r = round(65535.0 / (double)colormap_length);
        uint16_t v = r >= 65535.0 ? 65535 : (r > 0 ? (uint16_t)r : 0);
        min_value = v > 1 ? v : 1;
    }
    float u16_span = (float)(65535 - min_value);
    for (uint64_t i = 0; i < height; i++) {
        uint64_t i_freq = i0 + i;
        for (uint64_t j = 0; j < width; j++) {
            uint16_t px = 0;
            if (i_freq < B) {
                volatile float zero_to_one = (spec[j * B + i_freq] - dB_min) / dB_span;
                volatile float prod = zero_to_one * u16_span; /* no FMA contraction: Rust does not fuse */
                float v = prod + (float)min_value;
                float r = roundf(v);
                /* f32::clamp(0, 65535) then `as u16` (NaN -> 0) */
                if (r != r) px = 0;
                else if (r < 0.f) px = 0;
                else if (r > 65535.f) px = 65535;
                else px = (uint16_t)r;
            }
            out[i * width + j] = px;
        }
    }
}

/* a12 + a15: global min/max clamp rule  (src-tauri/src/core/mod.rs:168-180) */
ORC_API void orc_clamp_minmax(float mn, float mx, float dB_range, float *min_dB, float *max_dB) {
    float mxx = fminf(mx, 0.f);
    float mnn = fmaxf(mn, mxx - dB_range);
    *min_dB = mnn; *max_dB = mxx;
}

/* ------------------------------------------------------------------------- */
/* a4-a10  calc_spec  (src-tauri/src/core/spectrogram.rs:187-212)              */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t hop, win, n_fft, n_freq, n_mel; /* n_mel == 0 -> Linear */
    float *window;  /* [win] f32, normalised */
    float *mel_fb;  /* (n_freq, n_mel) f32 or NULL */
    orc_plan_f32 *p32;
    orc_plan_f64 *p64;
} orc_analyzer;

/* freq_scale 0 Linear / 1 Mel; n_mel_req == 0 -> reference default rule */
ORC_API orc_analyzer *orc_analyzer_new(uint32_t sr, double win_ms, uint32_t t_overlap,
                                       uint32_t f_overlap, int freq_scale, uint64_t n_mel_req) {
    orc_analyzer *a = (orc_analyzer *)calloc(1, sizeof(*a));
    orc_framing_params(win_ms, sr, t_overlap, f_overlap, &a->hop, &a->win, &a->n_fft);
    a->n_freq = a->n_fft / 2 + 1;
    a->window = (float *)malloc(sizeof(float) * (a->win ? a->win : 1));
    orc_normalized_hann_f32(a->win, a->n_fft, a->window);
    if (freq_scale == 1) {
        if (n_mel_req == 0) {
            uint64_t cap = orc_mel_fb_default_f32(sr, a->n_fft, NULL);
            a->n_mel = cap;
            a->mel_fb = (float *)malloc(sizeof(float) * a->n_freq * cap);
            orc_mel_fb_default_f32(sr, a->n_fft, a->mel_fb);
        } else {
            a->n_mel = n_mel_req;
            a->mel_fb = (float *)malloc(sizeof(float) * a->n_freq * n_mel_req);
            orc_mel_fb_f32(sr, a->n_fft, n_mel_req, 0.f, 0, 0.f, 1, a->mel_fb);
        }
    }
    a->p32 = orc_plan_new_f32(a->n_fft);
    a->p64 = orc_plan_new_f64(a->n_fft);
    return a;
}
ORC_API void orc_analyzer_free(orc_analyzer *a) {
    if (!a) return;
    free(a->window); free(a->mel_fb);
    orc_plan_free_f32(a->p32); orc_plan_free_f64(a->p64);
    free(a);
}
ORC_API void orc_analyzer_dims(const orc_analyzer *a, uint64_t *hop, uint64_t *win, uint64_t *n_fft,
                               uint64_t *n_bins) {
    *hop = a->hop; *win = a->win; *n_fft = a->n_fft;
    *n_bins = a->n_mel ? a->n_mel : a->n_freq;
}
ORC_API const float *orc_analyzer_window(const orc_analyzer *a) { return a->window; }
ORC_API const float *orc_analyzer_mel_fb(const orc_analyzer *a) { return a->mel_fb; }

static inline float orc_sample_reflect(const float *x, uint64_t n, int64_t s) {
    return x[orc_reflect_index(s, (int64_t)n)];
}

/* One frame, f32 reference-like.  mag: [n_freq]; out: [n_bins] dB. */
static void orc_frame_f32(const orc_analyzer *a, const float *x, uint64_t n, uint64_t t, float *buf,
                          float *re, float *im, float *work, float *mag, float *out,
                          float *stft_re, float *stft_im) {
    const uint64_t W = a->win, NF = a->n_fft;
    const uint64_t pad_l = (NF - W) / 2; /* stft.rs:36-37 */
    memset(buf, 0, sizeof(float) * NF);
    int64_t s0 = (int64_t)(t * a->hop) - (int64_t)(W / 2);
    if (s0 >= 0 && (uint64_t)s0 + W <= n) { /* mid frames: a plain slice (stft.rs:83-85) */
        const float *xs = x + s0;
        for (uint64_t i = 0; i < W; i++) buf[pad_l + i] = xs[i] * a->window[i];
    } else {
        for (uint64_t i = 0; i < W; i++) buf[pad_l + i] = orc_sample_reflect(x, n, s0 + (int64_t)i) * a->window[i];
    }
    orc_rfft_f32(a->p32, buf, re, im, work);
    if (stft_re) { memcpy(stft_re, re, sizeof(float) * a->n_freq); memcpy(stft_im, im, sizeof(float) * a->n_freq); }
    for (uint64_t k = 0; k < a->n_freq; k++) mag[k] = hypotf(re[k], im[k]); /* Complex::norm */
    (void)out;
}

/* dense (T,F)x(F,M) f32 product, row block at a time -- the reference's `linspec.dot(&mel_fb)`
   (spectrogram.rs:207) is an sgemm over ALL F*M weights, zeros included (OpenBLAS via ndarray's
   `blas` feature).  Register-blocked 4 rows x 16 columns so that the CPU baseline is not
   handicapped by a naive loop; FMA where the CPU has it, like OpenBLAS' kernels. */
#ifdef __FMA__
#define ORC_MAC(acc, a, b) (acc) = __builtin_fmaf((a), (b), (acc))
#else
#define ORC_MAC(acc, a, b) (acc) += (a) * (b)
#endif
static void orc_dense_mel_rows_f32(const float *mag, uint64_t rows, uint64_t F, const float *fb,
                                   uint64_t M, float *out) {
    const uint64_t M16 = M & ~(uint64_t)15;
    uint64_t r = 0;
    for (; r + 4 <= rows; r += 4) {
        const float *a0 = mag + (r + 0) * F, *a1 = mag + (r + 1) * F, *a2 = mag + (r + 2) * F, *a3 = mag + (r + 3) * F;
        for (uint64_t m0 = 0; m0 < M16; m0 += 16) {
            float c0[16] = {0}, c1[16] = {0}, c2[16] = {0}, c3[16] = {0};
            for (uint64_t k = 0; k < F; k++) {
                const float *w = fb + k * M + m0;
                const float v0 = a0[k], v1 = a1[k], v2 = a2[k], v3 = a3[k];
#pragma omp simd
                for (int j = 0; j < 16; j++) {
                    ORC_MAC(c0[j], v0, w[j]); ORC_MAC(c1[j], v1, w[j]);
                    ORC_MAC(c2[j], v2, w[j]); ORC_MAC(c3[j], v3, w[j]);
                }
            }
            for (int j = 0; j < 16; j++) {
                out[(r + 0) * M + m0 + j] = c0[j]; out[(r + 1) * M + m0 + j] = c1[j];
                out[(r + 2) * M + m0 + j] = c2[j]; out[(r + 3) * M + m0 + j] = c3[j];
            }
        }
        for (uint64_t m = M16; m < M; m++) {
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            for (uint64_t k = 0; k < F; k++) {
                const float w = fb[k * M + m];
                ORC_MAC(c0, a0[k], w); ORC_MAC(c1, a1[k], w); ORC_MAC(c2, a2[k], w); ORC_MAC(c3, a3[k], w);
            }
            out[(r + 0) * M + m] = c0; out[(r + 1) * M + m] = c1; out[(r + 2) * M + m] = c2; out[(r + 3) * M + m] = c3;
        }
    }
    for (; r < rows; r++) {
        float *o = out + r * M;
        for (uint64_t m = 0; m < M; m++) o[m] = 0.f;
        const float *mr = mag + r * F;
        for (uint64_t k = 0; k < F; k++) {
            const float v = mr[k];
            const float *w = fb + k * M;
#pragma omp simd
            for (uint64_t m = 0; m < M; m++) ORC_MAC(o[m], v, w[m]);
        }
    }
}

/* calc_spec, f32 reference-like.  out: (T, n_bins) dB.  Optional stft_re/stft_im (T, n_freq).
   n_threads <= 1: serial (the reference's `parallel == false` leg); otherwise frames are split
   across OpenMP threads (the `parallel` leg, stft.rs:100-113). Returns T. */
ORC_API uint64_t orc_calc_spec_f32(const orc_analyzer *a, const float *x, uint64_t n, float *out,
                                   float *stft_re, float *stft_im, int n_threads) {
    const uint64_t T = orc_n_frames(n, a->win, a->hop);
    if (!out) return T;
    const uint64_t F = a->n_freq, NF = a->n_fft, M = a->n_mel;
    const uint64_t NB = M ? M : F;
    const uint64_t BLK = 64;
    (void)n_threads;
#pragma omp parallel num_threads(n_threads > 1 ? n_threads : 1)
    {
        float *buf = (float *)malloc(sizeof(float) * NF);
        float *re = (float *)malloc(sizeof(float) * F);
        float *im = (float *)malloc(sizeof(float) * F);
        float *work = (float *)malloc(sizeof(float) * 4 * (NF / 2 ? NF / 2 : 1));
        float *mag = (float *)malloc(sizeof(float) * F * BLK);
#pragma omp for schedule(dynamic, 1)
        for (uint64_t t0 = 0; t0 < T; t0 += BLK) {
            uint64_t rows = T - t0 < BLK ? T - t0 : BLK;
            for (uint64_t r = 0; r < rows; r++)
                orc_frame_f32(a, x, n, t0 + r, buf, re, im, work, mag + r * F, NULL,
                              stft_re ? stft_re + (t0 + r) * F : NULL,
                              stft_im ? stft_im + (t0 + r) * F : NULL);
            float *o = out + t0 * NB;
            if (M) orc_dense_mel_rows_f32(mag, rows, F, a->mel_fb, M, o);
            else memcpy(o, mag, sizeof(float) * rows * F);
            /* dB_from_amp_inplace_default: log pass then *20 pass (decibel.rs:198-202) */
            for (uint64_t i = 0; i < rows * NB; i++) {
                float v = o[i];
                float y;
                if (v != v || signbit(v)) y = NAN;
                else if (v > 0.f) y = log10f(v) - 0.f;
                else y = -INFINITY;
                o[i] = y;
            }
            for (uint64_t i = 0; i < rows * NB; i++) o[i] *= 20.f;
        }
        free(buf); free(re); free(im); free(work); free(mag);
    }
    return T;
}

/* calc_spec, f64 truth.  out_db: (T, n_bins) double dB; optional out_amp (T, n_bins) linear
   amplitude (|X| or mel-projected |X|) and out_pow (T, n_freq) |X|^2. */
ORC_API uint64_t orc_calc_spec_f64(const orc_analyzer *a, const float *x, uint64_t n, double *out_db,
                                   double *out_amp, double *out_pow, int n_threads) {
    const uint64_t T = orc_n_frames(n, a->win, a->hop);
    if (!out_db && !out_amp && !out_pow) return T;
    const uint64_t F = a->n_freq, NF = a->n_fft, M = a->n_mel, W = a->win;
    const uint64_t NB = M ? M : F;
    const uint64_t pad_l = (NF - W) / 2;
    (void)n_threads;
#pragma omp parallel num_threads(n_threads > 1 ? n_threads : 1)
    {
        double *buf = (double *)malloc(sizeof(double) * NF);
        double *re = (double *)malloc(sizeof(double) * F);
        double *im = (double *)malloc(sizeof(double) * F);
        double *work = (double *)malloc(sizeof(double) * 4 * (NF / 2 ? NF / 2 : 1));
        double *mag = (double *)malloc(sizeof(double) * F);
        double *mel = (double *)malloc(sizeof(double) * (M ? M : 1));
#pragma omp for schedule(dynamic, 16)
        for (uint64_t t = 0; t < T; t++) {
            memset(buf, 0, sizeof(double) * NF);
            int64_t s0 = (int64_t)(t * a->hop) - (int64_t)(W / 2);
            for (uint64_t i = 0; i < W; i++)
                buf[pad_l + i] = (double)orc_sample_reflect(x, n, s0 + (int64_t)i) * (double)a->window[i];
            orc_rfft_f64(a->p64, buf, re, im, work);
            for (uint64_t k = 0; k < F; k++) {
                double p = re[k] * re[k] + im[k] * im[k];
                if (out_pow) out_pow[t * F + k] = p;
                mag[k] = hypot(re[k], im[k]);
            }
            const double *v = mag;
            if (M) {
                for (uint64_t m = 0; m < M; m++) mel[m] = 0.0;
                for (uint64_t k = 0; k < F; k++) {
                    const float *w = a->mel_fb + k * M;
                    for (uint64_t m = 0; m < M; m++)
                        if (w[m] != 0.f) mel[m] += mag[k] * (double)w[m];
                }
                v = mel;
            }
            for (uint64_t b = 0; b < NB; b++) {
                if (out_amp) out_amp[t * NB + b] = v[b];
                if (out_db) out_db[t * NB + b] = orc_dB_amp_default_f64(v[b]);
            }
        }
        free(buf); free(re); free(im); free(work); free(mag); free(mel);
    }
    return T;
}

/* perform_stft restated literally with the three-piece framing (for stft_works etc.):
   out (T, n_fft/2+1) complex as separate re/im. window == NULL -> default normalised Hann. */
ORC_API uint64_t orc_perform_stft_f32(const float *x, uint64_t n, uint64_t win, uint64_t hop,
                                      uint64_t n_fft, const float *window, float *out_re, float *out_im) {
    uint64_t T = orc_stft_frames_f32(x, n, win, hop, NULL);
    if (!out_re) return T;
    float *frames = (float *)malloc(sizeof(float) * ((T * win) > 0 ? T * win : 1));
    orc_stft_frames_f32(x, n, win, hop, frames);
    float *w = (float *)malloc(sizeof(float) * (win ? win : 1));
    if (window) memcpy(w, window, sizeof(float) * win);
    else orc_normalized_hann_f32(win, n_fft, w);
    uint64_t F = n_fft / 2 + 1;
    uint64_t pad_l = (n_fft - win) / 2;
    orc_plan_f32 *p = orc_plan_new_f32(n_fft);
    float *buf = (float *)malloc(sizeof(float) * n_fft);
    float *work = (float *)malloc(sizeof(float) * 4 * (n_fft / 2 ? n_fft / 2 : 1));
    for (uint64_t t = 0; t < T; t++) {
        memset(buf, 0, sizeof(float) * n_fft);
        for (uint64_t i = 0; i < win; i++) buf[pad_l + i] = frames[t * win + i] * w[i]; /* stft.rs:139-142 */
        orc_rfft_f32(p, buf, out_re + t * F, out_im + t * F, work);
    }
    free(buf); free(work); free(w); free(frames);
    orc_plan_free_f32(p);
    return T;
}

/* ------------------------------------------------------------------------- */
/* a11/a15 batch driver = TrackManager::update_specs + update_spec_imgs        */
/*         (src-tauri/src/core/mod.rs:137-230) -- used as the timed CPU baseline */
/* ------------------------------------------------------------------------- */
/* All channels share (sr, setting).  specs[c]: (T_c, NB) dB f32 (caller-allocated),
   imgs[c]: (NB, T_c) u16 or NULL to skip the image pass.  Threading mirrors mod.rs:152:
   across channels when n_ch >= n_threads, else across frames inside each channel. */
ORC_API void orc_update_specs_and_imgs(const orc_analyzer *a, const float *const *pcm,
                                       const uint64_t *lens, uint64_t n_ch, float **specs,
                                       uint16_t **imgs, float dB_range, uint32_t colormap_length,
                                       int n_threads, float *min_dB, float *max_dB) {
    const uint64_t NB = a->n_mel ? a->n_mel : a->n_freq;
    if (n_threads < 1) n_threads = 1;
    int across_tracks = (int)n_ch >= n_threads; /* parallel = n < threads (mod.rs:152) */
    if (across_tracks) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
        for (uint64_t c = 0; c < n_ch; c++) orc_calc_spec_f32(a, pcm[c], lens[c], specs[c], NULL, NULL, 1);
    } else {
        for (uint64_t c = 0; c < n_ch; c++) orc_calc_spec_f32(a, pcm[c], lens[c], specs[c], NULL, NULL, n_threads);
    }
    float gmn = INFINITY, gmx = -INFINITY;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(min : gmn) reduction(max : gmx)
    for (uint64_t c = 0; c < n_ch; c++) {
        float mn, mx;
        orc_find_min_max_f32(specs[c], orc_n_frames(lens[c], a->win, a->hop) * NB, &mn, &mx);
        gmn = fminf(gmn, mn); gmx = fmaxf(gmx, mx);
    }
    orc_clamp_minmax(gmn, gmx, dB_range, min_dB, max_dB);
    if (imgs) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
        for (uint64_t c = 0; c < n_ch; c++)
            if (imgs[c])
                orc_spec_to_img(specs[c], orc_n_frames(lens[c], a->win, a->hop), NB, 0, NB, *min_dB,
                                *max_dB, 1, colormap_length, imgs[c]);
    }
}

/* ------------------------------------------------------------------------- */
/* Synthetic PCM twin of thb_synth_pcm (integer arithmetic; bit-identical to  */
/* thesia_b200/synth.py and to the device generator) -- feeds the CPU baseline */
/* ------------------------------------------------------------------------- */
static inline int orc_para_sine(uint32_t phase) {
    uint32_t x = (phase >> 15) & 0xffffu;
    int y = (int)(((uint64_t)x * (65536ull - x)) >> 14);
    return (phase >> 31) ? -y : y;
}
static inline uint32_t orc_mix32(uint32_t h) {
    h ^= h >> 15; h *= 0x85ebca77u;
    h ^= h >> 13; h *= 0xc2b2ae3du;
    h ^= h >> 16;
    return h;
}
static inline int orc_synth_base(uint64_t n, uint64_t len, uint32_t sr, uint32_t track, uint32_t flags) {
    uint64_t f0_mhz = 55000ull + 13750ull * (track % 61u);
    uint64_t inc0 = (f0_mhz << 32) / (1000ull * sr);
    uint32_t ph0 = (uint32_t)(n * inc0);
    uint64_t inc_a = (50ull << 32) / sr;
    uint64_t dinc = 1932735283ull - inc_a;
    uint64_t n2 = n * n, two_len = 2ull * len;
    uint64_t t_int = n2 / two_len, t_rem = n2 % two_len;
    uint32_t ph1 = (uint32_t)(inc_a * n + dinc * t_int + (dinc * t_rem) / two_len);
    uint32_t h = orc_mix32((uint32_t)n * 0x9e3779b1u + track * 0x7f4a7c15u + 0x7e51au);
    int a0 = (8192 * orc_para_sine(ph0)) >> 16;
    int a1 = (3277 * orc_para_sine(ph1)) >> 16;
    int nz = (((int)(h >> 16) - 32768) * 1638) >> 15;
    int v = a0 + a1 + nz;
    if ((flags & 2u) && n >= sr && n < 2ull * sr) v = 0;
    if (flags & 1u) v *= 32;
    return v;
}
ORC_API void orc_synth_pcm(float *out, uint64_t len, uint32_t sr, uint32_t track, uint32_t channel,
                           uint32_t flags, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads > 1 ? n_threads : 1)
    for (uint64_t n = 0; n < len; n++) {
        int v;
        if (channel == 0) v = orc_synth_base(n, len, sr, track, flags);
        else v = n >= 7 ? (int)(((int64_t)orc_synth_base(n - 7, len, sr, track, flags) * 26214) >> 15) : 0;
        out[n] = (float)v / 32768.0f;
    }
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
