// thb_kernels.cuh -- launch interface between the C-ABI layer (thb_api.cu) and the sm_100a
// kernels (thb_stft.cu, thb_image.cu, thb_envelope.cu).  Device pointers only.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>

namespace thb {

// One channel (or one frame-range shard of a channel) as the STFT kernels see it.
struct TrackDesc {
    const float *pcm;        // device; pcm[0] is sample `pcm_offset` of the file (int16_t samples when pcm_i16)
    long long pcm_offset;
    long long slice_len;     // samples available at pcm
    long long full_len;      // samples in the whole file (reflect happens at 0 and full_len-1)
    long long frame_begin;   // first frame of this shard
    long long n_frames;      // frames to compute
    float *out;              // (n_frames, n_bins) row-major dB
    float *minmax;           // 2 floats: {max, -min} of this channel (atomic-max accumulated)
    int pcm_i16;             // 1: 16-bit PCM, sample value = s / 32768 (exact)
    int pad_;
};

// Everything that depends only on (sr, win, n_fft, n_mel): uploaded once per SrWinNfft key.
struct PlanDev {
    int hop, win, n_fft, nc;          // nc = n_fft / 2
    int pad_left;                     // (n_fft - win) / 2  (stft.rs:36)
    int n_freq, n_bins, n_mel;        // n_mel == 0 -> linear
    int n_pass;
    int radix_log2[8];                // DIF pass radices (log2), product = nc
    const float *window;              // [win]
    const float2 *twiddle;            // [n_fft] exp(-2 pi i t / n_fft)
    const uint32_t *mel_k0;           // [n_mel]
    const uint32_t *mel_ptr;          // [n_mel + 1]
    const float *mel_w;               // [nnz]
    int max_band_len;
    int mel_nnz;
    // warp-register paths: n_fft == 2048 (thb_stft_fast.cu / thb_stft_pair.cu) and n_fft == 1024 / 512 (thb_stft_warp.cu),
    // R1 = n_fft / 64; null otherwise
    const float *fast_wpad;           // [n_fft] 0.5 * window centred in the FFT buffer, zeros outside
    const float2 *fast_tw;            // [(R1-1)*32] W_(n_fft/2)^(lane*k1) then [16*32] split twiddles
    // warp schedule of the sparse mel product (thb_host.hpp MelItems::blob); mel only, n_fft == 2048
    int mi_words;                     // size of mi_blob in 32-bit words
    int mi_groups, mi_min_start, mi_max_reach;
    int mi_direct;                    // 1: the blob holds the band-major schedule (mel_direct), 0: the bin-major one
    const uint32_t *mi_blob;
    // n_fft == 16384 two-frame path (thb_stft_big.cu); null otherwise
    const float *big_wpad;            // [16384] 0.5 * window centred in the FFT buffer, zeros outside
    const float2 *big_tw;             // [31][16] W_512^(n2 k1); [32][16] W_8192^(n3 k1); [16][16] W_256^(n3 k2)
    // mel: band-major pieces of <= 32 bins, 32 consecutive pieces = one warp-wide group
    const uint32_t *big_pieces;       // [32 G] first bin of every piece, then [G] {steps, offset into big_w}
    const float *big_w;               // per group [steps][32 lanes], zero padded
    const uint32_t *big_pptr;         // [n_mel + 1] pieces of band m: big_pptr[m] .. big_pptr[m + 1]
    int big_n_pieces;
    int load_all_rows;                // thb_stft_warp.cu, A/B switch (THB_WARP_ALLROWS=1): also load the rows of the FFT buffer that hold no window tap
};

// Tiles (tile_frames consecutive frames of one descriptor) that the frame-pair kernel could not finish
// exactly and the scalar kernel redoes: items[i] = {descriptor index, tile index}; flags[] (one word per
// (descriptor, tile), zeroed before the launch) keeps a tile from being listed twice.
struct RescueList {
    uint2 *items;
    unsigned *count;
    unsigned *flags;
    unsigned capacity;
    unsigned tile_frames;
    unsigned tiles_per_track;
};

struct ImgDesc {
    const float *spec;   // (T, B) dB
    uint16_t *img;       // (H, pitch) u16
    long long T;
    int B;
    int i0, H;           // rows i0 .. i0+H of the bins axis
    long long pitch;     // elements between image rows
};

struct EnvDesc {
    const float *pcm;    // device, whole channel
    long long len;
    uint8_t *out;        // device, tiles of this level concatenated in wire format
};

// One channel of thb_apply_gain (thb_gain.cu)
struct GainDesc {
    const void *in;      // device; f32 samples, or int16_t when pcm_i16
    float *out;          // device, len floats (may alias `in` for f32)
    float *before;       // device or null: gain * x before clipping (Clip mode)
    long long len;
    float gain;
    int group;           // index of the track this channel belongs to (ReduceGlobalLevel peaks are per track)
    int pcm_i16;
    int pad_;
};
struct GainOut {
    float sum_squares;               // of the output channel
    unsigned abs_max_bits;           // bit pattern of the output's abs max (>= 0)
    unsigned before_max_bits;        // bit pattern of max |gain * x| (Clip mode)
    unsigned pad_;
    unsigned long long reduction_cnt;  // samples with |gain * x| > 1 (Clip mode)
};

// One spectrogram tile (thb_tile.cu): the two axes of the resize (ResizeAxis, thb_host.hpp) and its buffers
struct TileDesc {
    const uint16_t *img;      // (H, pitch) u16, retained image
    unsigned long long pitch;
    const unsigned *x_start, *x_size, *y_start, *y_size;
    const int *wx, *wy;       // tap-major: wx[i * width + x], wy[i * height + y]
    unsigned width, height;   // of the tile
    unsigned y_first, tmp_h;  // rows [y_first, y_first + tmp_h) of the image feed the vertical pass
    unsigned px, py;          // fixed-point precisions
    uint16_t *tmp;            // (tmp_h, width)
    uint8_t *out;             // (height, width) RGBA, rows reversed
    unsigned identity;        // both axes copy (one tap of weight 2^precision, consecutive starts): level 0 in x and y
    unsigned x_first;         // identity: first image column of the tile
};

// sample `idx` of a channel's slice as f32: 16-bit PCM converts as s / 32768 (exact)
__device__ __forceinline__ float pcm_sample(const TrackDesc &d, long long idx) {
    if (d.pcm_i16) return static_cast<float>(__ldg(reinterpret_cast<const short *>(d.pcm) + idx)) * 3.0517578125e-05f;
    return __ldg(d.pcm + idx);
}

// Launch `kern` so that it may be scheduled under the tail of the kernel before it in the stream (programmatic stream
// serialisation); the kernel's first statement is pdl_wait().  What it buys is the launch latency between the small
// dependent kernels of a step (rescue list -> range exchange -> quantiser): a few microseconds each, which is what a
// 250 us step of a frame-range shard at N = 8 is made of.  THB_PDL=0 launches the ordinary way.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    static const bool on = !(getenv("THB_PDL") && atoi(getenv("THB_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- launchers (all asynchronous on `st`) ----
// generic shared-memory path, any power-of-two n_fft in [4, 32768]
cudaError_t launch_stft_generic(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks,
                                long long max_frames, cudaStream_t st);
// warp-per-frame register path for n_fft == 2048 (nc == 1024)
bool stft_fast_supported(const PlanDev &plan);
cudaError_t launch_stft_fast(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks,
                             long long max_frames, int sm_count, cudaStream_t st);

// two frames per CTA in packed f32x2 arithmetic, n_fft == 16384 (handles every frame itself: edges, 16-bit PCM)
bool stft_big_supported(const PlanDev &plan);
// float2 slots of its shared-memory buffer (compact magnitudes + mel partial sums must fit)
int stft_big_buffer_slots(int n_fft);
cudaError_t launch_stft_big(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, long long max_frames,
                            int sm_count, cudaStream_t st);

// two frames per warp in packed f32x2 arithmetic, n_fft == 2048
bool stft_pair_supported(const PlanDev &plan);
// (every descriptor of one launch holds the same PCM format: pcm_i16 says which; `unaligned`: f32 frames need not
// start on an 8-byte boundary -- odd hops, odd channel starts -- and are read with 4-byte loads)
cudaError_t launch_stft_pair(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, RescueList rescue,
                             bool pcm_i16, bool unaligned, int sm_count, cudaStream_t st);
// frames per work item of the frame-pair kernel (a multiple of twice its warps per CTA)
int stft_pair_tile_frames();

// n_fft == 1024 / 512 in the same warp-register design (thb_stft_warp.cu): the packed kernel over the interior frames of
// every descriptor (any count; `unaligned` as above), its scalar twin over whole descriptors (file edges, leftovers)
// and over the rescue list; both agree bit for bit on every frame
bool stft_warp_supported(const PlanDev &plan);
int stft_warp_tile_frames(const PlanDev &plan);
cudaError_t launch_stft_warp_packed(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, RescueList rescue,
                                    bool pcm_i16, bool unaligned, int sm_count, cudaStream_t st);
cudaError_t launch_stft_warp_scalar(const PlanDev &plan, const TrackDesc *d_tracks, int n_tracks, long long max_frames,
                                    int sm_count, cudaStream_t st);
cudaError_t launch_stft_warp_list(const PlanDev &plan, const TrackDesc *d_tracks, RescueList rescue, int sm_count,
                                  cudaStream_t st);
// the scalar kernel over the tiles on a rescue list (persistent grid; a no-op when the list is empty)
cudaError_t launch_stft_fast_list(const PlanDev &plan, const TrackDesc *d_tracks, RescueList rescue,
                                  int sm_count, cudaStream_t st);

cudaError_t launch_minmax_init(float *d_slots, int n, cudaStream_t st);
// reset the {max, -min} slot of every channel in a descriptor array
cudaError_t launch_minmax_init_tracks(const TrackDesc *d_tracks, int n, cudaStream_t st);
// {max, -min} of an arbitrary f32 array accumulated into slot[0..1] (find_min_max, simd.rs:14-36)
cudaError_t launch_minmax_array(const float *d_x, unsigned long long n, float *d_slot, int sm_count, cudaStream_t st);
// d_send[0..1] = max over live slots of {max, -min}
cudaError_t launch_minmax_reduce(const float *d_slots, int n_slots, float *d_send, cudaStream_t st);
// reduce + exchange over peer memory + finalize in one launch (thb_image.cu); d_peers[r] / d_mine: the exchange slots
// of rank r / of this rank, 2 x kExchangeMaxRanks float4 each
constexpr int kExchangeMaxRanks = 16;
cudaError_t launch_minmax_exchange(const float *d_slots, int n_slots, float4 *const *d_peers, float4 *d_mine, int n_ranks, int rank,
                                   unsigned seq, float dB_range, float *d_range, float *d_send, unsigned *d_fail, cudaStream_t st);
// d_range = {min_dB, max_dB} after the clamp rules of mod.rs:179-180
cudaError_t launch_minmax_finalize(const float *d_send, float dB_range, float *d_range, cudaStream_t st);

// tile_mode 0: 32 x 64 tiles, any layout.  1: 128 x 128 tiles (image pitch even, 4-byte aligned rows).
// 2: as 1 with 16-byte loads (every descriptor: spec 16-byte aligned, B % 4 == 0, i0 % 4 == 0).
cudaError_t launch_spec_to_img(const ImgDesc *d_descs, int n, long long max_T, int max_H,
                               const float *d_range, uint32_t colormap_length, int tile_mode, cudaStream_t st);

// tiles [tile_begin, tile_begin + tile_count) of every channel; tile_count == 0 -> to the end
cudaError_t launch_envelope(const EnvDesc *d_descs, int n, long long max_len, uint32_t level,
                            uint64_t revision, uint32_t tile_begin, uint32_t tile_count,
                            cudaStream_t st);

// per-channel sum of squares / absolute maximum (thb_stats.cu); only pcm, slice_len and pcm_i16 of a descriptor are read
long long stats_chunks(long long max_len);
cudaError_t launch_channel_stats(const TrackDesc *d_descs, int n, long long max_len, double *d_part_ss, float *d_part_mx,
                                 float *d_out_ss, float *d_out_mx, cudaStream_t st);

// gain + guard clipping (thb_gain.cu).  mode 0 Clip, 1 ReduceGlobalLevel (d_group_peak from launch_gain_peak), 2 copy.
// d_part_ss: n * gain_chunks(max_len) doubles of scratch; d_outs: n results, zeroed by the caller.
long long gain_chunks(long long max_len);
cudaError_t launch_gain_peak(const GainDesc *d_descs, int n, long long max_len, unsigned *d_group_peak, cudaStream_t st);
cudaError_t launch_gain_apply(const GainDesc *d_descs, int n, long long max_len, int mode, const unsigned *d_group_peak,
                              double *d_part_ss, GainOut *d_outs, cudaStream_t st);

// spectrogram tiles (thb_tile.cu): grid y is sized for the largest tile of the batch, the others return early;
// n_identity = how many descriptors have `identity` set (the copy kernel / the two resample kernels are launched only
// when the batch holds tiles of their kind); spectrogram_tile_launches = the number of launches that makes
int spectrogram_tile_launches(int n, int n_identity);
cudaError_t launch_spectrogram_tiles(const TileDesc *d_descs, int n, int n_identity, unsigned max_w, unsigned max_h, unsigned max_tmp_h,
                                     const uchar4 *d_colormap, unsigned colors, cudaStream_t st);

cudaError_t launch_synth_pcm(float *d_out, unsigned long long len, uint32_t sr, uint32_t track,
                             uint32_t channel, uint32_t flags, cudaStream_t st);

}  // namespace thb
