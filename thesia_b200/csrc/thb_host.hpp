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
// order to agree bit for bit).  The bins are walked ONCE: bin k between the peaks of bands s - 1 and s ("segment" s)
// carries two weights, the rising one of band s and the falling one of band s - 1.  Every segment is cut into
// pieces of at most kMelPieceMax consecutive bins; the pieces are sorted by length and dealt 32 at a time to the
// lanes of a warp ("groups"), so that a group costs max-length steps.  Lane l of group g walks T[g] bins from
// start[g*32+l]; at step t it multiplies the bin by the weight pair w[w_index(g, t, l) + {0: rise, 1: fall}]
// (zero outside the piece) into two accumulators, which end up in slots g*32+l (rise) and n_slots + g*32+l (fall),
// n_slots = 32 n_groups.  Band m is the sum of the slots piece_ids[piece_ptr[m] .. piece_ptr[m+1]): the rise sums of
// segment m's pieces, then the fall sums of segment m+1's, each in ascending-bin order.  The kernels read those
// lists as a table: round r (bands 32 r ..) has gk[r] rows starting at row gbase[r]; row j holds one slot id per
// lane (goff[(gbase[r] + j) * 32 + lane]), lists shorter than gk[r] are padded with zero_slot, a slot nobody writes.
// Shared-memory banks: `start` is pulled back by up to 15 bins (zero weights) so that the 16 lanes of each half
// warp start on 16 different residues mod 16 -- their float2 loads then hit 16 different 8-byte bank pairs at every
// step -- and a piece of segment s sits in a lane congruent to s mod 16 when that lane is free, which spreads the
// gather of consecutive bands over the banks.  `valid` is false when the bank is not a chain of overlapping
// triangles (at most two adjacent bands per bin); such banks run on the generic kernel.
constexpr uint32_t kMelPieceMax = 15;
struct MelItems {
    bool valid = false;
    uint32_t n_groups = 0, n_mel = 0;
    std::vector<uint32_t> T;          // [n_groups], even
    std::vector<uint32_t> woff;       // [n_groups], offset into w
    std::vector<int32_t> start;       // [n_groups * 32]
    std::vector<float> w;             // float4 per lane and two steps: {rise, fall} of step 2i, then of step 2i+1
    std::vector<uint32_t> piece_ptr;  // [n_mel + 1]
    std::vector<uint32_t> piece_ids;  // slots
    std::vector<uint32_t> gk, gbase;  // [ceil(n_mel / 32)]
    std::vector<uint16_t> goff;       // [sum(gk) * 32]
    // the same gather lists for the n_fft == 2048 kernels: per round ceil(gk / 4) rows of 32 lanes x 4 BYTE offsets
    // (slot * 8: float2 slots of the frame-pair kernel; the scalar kernel halves them), padded with the zero slot
    std::vector<uint32_t> gk4, gbase4;  // [rounds]: rows of 4, first row
    std::vector<uint32_t> goff4;        // [sum(gk4) * 32 * 4]
    // Band-major alternative for banks of NARROW bands (the default banks: 257 - 404 bands of 2 - 50 bins): lane l of
    // round r owns band 32 r + l and walks its own bins, direct_L[r] steps (the longest band of the round; shorter ones
    // run on zero weights): no partial sums, no gather.  direct_w is step-major per round ([step][lane], coalesced).
    // use_direct says which schedule costs fewer instructions for this bank (cost model in thb_host.cpp).
    std::vector<uint32_t> direct_L, direct_woff;  // [rounds]
    std::vector<int32_t> direct_k0;               // [rounds * 32] first bin of the lane's band
    std::vector<float> direct_w;
    bool use_direct = false;
    uint32_t direct_reach = 0;                    // largest bin index the band-major walk reads
    uint32_t zero_slot = 0;
    size_t w_index(uint32_t g, uint32_t t, uint32_t lane) const {
        return woff[g] + static_cast<size_t>(t / 2) * 128 + 4 * static_cast<size_t>(lane) + 2 * (t & 1);
    }
    int32_t min_start = 0;
    uint32_t max_reach = 0;           // largest bin index any lane reads
    // one blob of 32-bit words for the device: header {n_groups, n_mel, off_groups, off_start, off_rounds,
    // off_goff, zero_slot, off_w} then the arrays (see blob())
    std::vector<uint32_t> blob() const;
};
// t_multiple: every group's step count is a multiple of it (the walks take two steps per float4 of weights; padding
// steps carry zero weights and add +0).  4 would spare the n_fft == 2048 walk its two-step tail but costs the default
// banks (10 - 11 groups of 2 - 12 steps) 12 - 14 extra steps per frame pair: measured on the host, not adopted.
// with_direct: also build the band-major schedule (n_fft <= 2048 kernels) and choose between the two
// direct_pairs: the band-major rounds come in pairs of equal step count (zero-weight padding; an odd count of rounds is
// completed with an empty round) for kernels that walk two rounds at a time (thb_stft_warp.cu: n_fft 1024 / 512)
MelItems mel_items(const MelBank &b, uint32_t t_multiple = 2, bool with_direct = false, bool direct_pairs = false);
float mel_from_hz(float hz);
float mel_to_hz(float mel);
// n_mel == 0 -> calc_mel_fb_default's rule
MelBank mel_bank(uint32_t sr, uint64_t n_fft, uint32_t n_mel);

void hz_range_to_idx(uint32_t freq_scale, float hz0, float hz1, uint32_t sr, uint64_t n_bins,
                     uint64_t *i0, uint64_t *i1);

// ---- spectrogram tiles (render_tiles.rs:281-393) ----
// geometry of a tile (render_tiles.rs:290-312)
struct TileGeometry {
    uint64_t lod_width = 0, lod_height = 0, origin_x = 0, origin_y = 0, width = 0, height = 0;
};
TileGeometry spectrogram_tile_geometry(uint64_t H, uint64_t W, uint32_t level_x, uint32_t level_y, uint32_t tile_x,
                                       uint32_t tile_y);
// One axis of the Lanczos3 convolution resize the reference delegates to fast_image_resize 6.0.0 (U16 pixels): the
// taps of output pixel o are input pixels start[o] .. start[o] + size[o], weights w_t[i * n + o] (tap-major, so that
// neighbouring output pixels read neighbouring words) in i32 fixed point with `precision` fractional bits.
struct ResizeAxis {
    uint32_t n = 0, window = 0, precision = 0;
    std::vector<uint32_t> start, size;
    std::vector<int32_t> w_t;
};
ResizeAxis resize_axis(uint32_t in_size, double in0, double in1, uint32_t out_size);

// exp(-2*pi*i*t/n_fft), t = 0..n_fft-1, computed in double and rounded once
std::vector<float> twiddle_table(uint64_t n_fft);

}  // namespace thb
