// thb_host.hpp -- host-side parameter arithmetic of the analysis path (no CUDA).
//
// Product code: mirrors, value for value, what the reference computes on the host before the
// per-frame hot loops: SpecSetting framing (spectrogram.rs:57-98), the normalised periodic Hann
// window (windows.rs:12-83), the mel filterbank and its default size rule
// (src-common/src/lib.rs:11-103) and hz_range_to_idx (lib.rs:135-159).  The tables are tiny
// (KBs), are cached per (sr, win, n_fft) like SpectrogramAnalyzer::prepare does
// (spectrogram.rs:116-154) and are uploaded once; the per-sample work is all on the device.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/thesia_b200.h"

namespace thb {

struct Framing {
    uint64_t hop = 0, win = 0, n_fft = 0;
};

Framing framing_params(const thb_setting &s, uint32_t sr);
uint64_t n_frames(uint64_t len, uint64_t win, uint64_t hop);
bool is_pow2(uint64_t x);

// calc_normalized_win(Hann, win, n_fft)
std::vector<float> normalized_hann(uint64_t win, uint64_t n_fft);

// Mel filterbank in the two shapes the library needs: the dense (F, M) matrix the reference
// multiplies by (only for thb_mel_fb / tests) and the band-major sparse form the kernel reads.
// Every weight equals the reference's f32 value; zeros are simply not stored.
struct MelBank {
    uint32_t n_freq = 0, n_mel = 0;
    std::vector<uint32_t> k0;   // [n_mel]   first FFT bin with a non-zero weight
    std::vector<uint32_t> ptr;  // [n_mel+1] offsets into w
    std::vector<float> w;       // weights of band m: w[ptr[m] .. ptr[m+1]) for bins k0[m]...
    std::vector<float> dense() const;  // (n_freq, n_mel) row-major
};
// Lane schedule of the sparse mel product for a warp: bands are taken 32 at a time (lane = band
// within the group); every lane walks `T[g]` consecutive bins starting at `start`, with its weights
// interleaved as w[woff[g] + 32 t + lane] (zero where the band has no weight).  `start` is pulled
// back by up to 31 bins so that the 32 lanes of a group always hit 32 different shared-memory
// banks (start mod 32 distinct); it can therefore be as low as -31.
struct MelSchedule {
    uint32_t n_groups = 0;
    std::vector<uint32_t> T;      // [n_groups], multiple of 4
    std::vector<uint32_t> woff;   // [n_groups]
    std::vector<int32_t> start;   // [n_groups * 32]
    std::vector<float> w;         // interleaved weights
    uint32_t max_reach = 0;       // largest bin index any lane reads
};
MelSchedule mel_schedule(const MelBank &b);
// Variant for the frame-pair kernel: magnitudes are float2 (two frames) read two bins at a time with
// 16-byte loads, so `start` is even and the bank rule applies per quarter warp to start/2 mod 8;
// weights are stored as float2 (steps 2t, 2t+1): w[woff[g] + 64 (t/2) + 2 lane + (t & 1)].
MelSchedule mel_schedule_pair(const MelBank &b);
float mel_from_hz(float hz);
float mel_to_hz(float mel);
// n_mel == 0 -> calc_mel_fb_default's rule
MelBank mel_bank(uint32_t sr, uint64_t n_fft, uint32_t n_mel);

void hz_range_to_idx(uint32_t freq_scale, float hz0, float hz1, uint32_t sr, uint64_t n_bins,
                     uint64_t *i0, uint64_t *i1);

// exp(-2*pi*i*t/n_fft), t = 0..n_fft-1, computed in double and rounded once
std::vector<float> twiddle_table(uint64_t n_fft);

}  // namespace thb
