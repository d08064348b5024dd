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
// Warp schedule of the sparse mel product, shared by the two n_fft == 2048 kernels (they must add in the same
// order to agree bit for bit).  Every band is cut into pieces of at most kMelPieceMax consecutive bins; the pieces
// ("items") are sorted by length and dealt 32 at a time to the lanes of a warp ("groups"), so that a group costs
// max-length steps and the sum over groups stays close to nnz / 32.  Lane l of group g walks T[g] bins from
// start[g*32+l] with weights w[woff[g] + 64 (t/2) + 2 l + (t&1)] (zero outside the piece) and leaves a partial sum in slot
// g*32+l; band m is the sum of its slots piece_ids[piece_ptr[m] .. piece_ptr[m+1]) in ascending-bin order.
// `start` is pulled back by a few bins (zero weights) so that the 16 lanes of each half warp hit every 8-byte
// shared-memory bank pair at most twice; it can be negative (>= -15).
constexpr uint32_t kMelPieceMax = 16;
struct MelItems {
    uint32_t n_groups = 0, n_mel = 0;
    std::vector<uint32_t> T;          // [n_groups], even
    std::vector<uint32_t> woff;       // [n_groups], offset into w
    std::vector<int32_t> start;       // [n_groups * 32]
    std::vector<float> w;             // interleaved weights
    std::vector<uint32_t> piece_ptr;  // [n_mel + 1]
    std::vector<uint32_t> piece_ids;  // slots, ascending bins within a band
    int32_t min_start = 0;
    uint32_t max_reach = 0;           // largest bin index any lane reads
    // one blob of 32-bit words for the device: header {n_groups, n_mel, off_T, off_woff, off_start, off_pptr,
    // off_pids, off_w} then the arrays; woff entries are made absolute word offsets into the blob
    std::vector<uint32_t> blob() const;
};
MelItems mel_items(const MelBank &b);
float mel_from_hz(float hz);
float mel_to_hz(float mel);
// n_mel == 0 -> calc_mel_fb_default's rule
MelBank mel_bank(uint32_t sr, uint64_t n_fft, uint32_t n_mel);

void hz_range_to_idx(uint32_t freq_scale, float hz0, float hz1, uint32_t sr, uint64_t n_bins,
                     uint64_t *i0, uint64_t *i1);

// exp(-2*pi*i*t/n_fft), t = 0..n_fft-1, computed in double and rounded once
std::vector<float> twiddle_table(uint64_t n_fft);

}  // namespace thb
