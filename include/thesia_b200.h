/* thesia_b200.h -- C ABI of the B200-native analysis hot path for Sytronik/thesia.
 *
 * The reference (Rust, src-tauri/) has no FFI for this path; everything is in-crate calls.
 * Each entry point below replaces one in-crate seam and cites it (paths relative to the
 * reference checkout).  INTEGRATION.md shows the Rust `extern "C"` block and the three call
 * sites a maintainer would change.
 *
 * Conventions:
 *  - plain pointers and sizes only; the caller owns every host buffer, the library never keeps a
 *    host pointer past the call.  Device-side results are retained under the (id, ch) key until
 *    thb_release().
 *  - every function returns 0 (THB_OK) or a negative thb_status; thb_last_error() gives text.
 *    The Rust shim maps non-zero to panic!/anyhow::Error (the reference unwrap()s here:
 *    stft.rs:47, release profile panic = "abort").
 *  - Every entry point is thread-safe.  The analysis calls (thb_spec_batch, thb_update_spec_imgs,
 *    thb_minmax_global, thb_spec_to_img, release, ...) take the context exclusively, as the reference
 *    serialises them on its "write-lock-worker" thread (interface.rs:12-56).  The tile readers
 *    (thb_waveform_tile, thb_spectrogram_tile, thb_spectrogram_tile_batch) take it SHARED and each call
 *    runs on its own stream: concurrent tile calls overlap on the device, as the reference serves tiles
 *    from concurrent IPC threads under read locks (lib.rs:343-389).
 *  - `pcm` pointers may be HOST or DEVICE memory (detected with cudaPointerGetAttributes).
 *    Host memory from thb_host_alloc() is pinned and copies from it are asynchronous.
 *  - There is no CPU fallback: without a CUDA device thb_ctx_create() fails with THB_ERR_CUDA.
 *    The functions in the "host parameter" group are pure host arithmetic and need no device.
 */
#ifndef THESIA_B200_H
#define THESIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THB_ABI_VERSION 5

typedef enum thb_status {
    THB_OK = 0,
    THB_ERR_INVALID = -1,      /* bad argument (NULL, len < 2, hop == 0, ...) */
    THB_ERR_UNSUPPORTED = -2,  /* n_fft not a power of two or > 32768 */
    THB_ERR_CUDA = -3,         /* CUDA runtime error / no device */
    THB_ERR_NOMEM = -4,
    THB_ERR_NOT_FOUND = -5,    /* unknown (id, ch) */
    THB_ERR_NCCL = -6,
    THB_ERR_SMALL_BUFFER = -7, /* caller buffer too small; *written / dims tell the need */
    THB_ERR_INTERNAL = -8      /* a self-check of the library failed (a bug) */
} thb_status;

typedef struct thb_ctx thb_ctx;

/* FreqScale (src-common/src/lib.rs:106-110) */
#define THB_FREQ_LINEAR 0u
#define THB_FREQ_MEL 1u

/* SpecSetting (src-tauri/src/core/spectrogram.rs:30-38) + the n_mel of calc_mel_fb
 * (src-common/src/lib.rs:46-53); n_mel == 0 selects calc_mel_fb_default's rule (lib.rs:91-103),
 * which is what TrackManager always uses. */
typedef struct thb_setting {
    double win_ms;
    uint32_t t_overlap;
    uint32_t f_overlap;
    uint32_t freq_scale; /* THB_FREQ_LINEAR | THB_FREQ_MEL */
    uint32_t n_mel;      /* 0 = reference default rule */
} thb_setting;

/* One (id, ch) channel = `tracklist[id].channel(ch)` (src-tauri/src/core/track.rs:92-94) with its
 * sample rate.  full_len .. frame_count shard a long file by FRAME RANGE across GPUs: `pcm` then
 * points at sample `pcm_offset` of a file of `full_len` samples and only frames
 * [frame_begin, frame_begin + frame_count) are computed; reflect padding is applied against the
 * true file ends.  All four zero = the whole channel. */
typedef struct thb_track {
    const void *pcm;  /* host or device; f32 samples, or i16 when pcm_format == THB_PCM_I16 */
    uint64_t len;     /* samples available at pcm */
    uint64_t id;
    uint32_t ch;
    uint32_t sr;
    uint64_t full_len;
    uint64_t pcm_offset;
    uint64_t frame_begin;
    uint64_t frame_count; /* 0 = all frames from frame_begin */
    uint32_t pcm_format;  /* THB_PCM_F32 (0, what Audio.wavs holds, audio.rs:24-30) | THB_PCM_I16 */
    uint32_t reserved;    /* 0 */
} thb_track;

/* PCM ingest formats.  THB_PCM_I16 hands over 16-bit PCM as the decoder produced it: the kernels convert on
 * the fly with the decoder's own rule, sample / 32768 (audio.rs:262-439 via symphonia's i16 -> f32; exact in
 * f32), so the results equal those of the f32 channel bit for bit while the host->device copy and the HBM read
 * are half the size. */
#define THB_PCM_F32 0u
#define THB_PCM_I16 1u

/* Per-channel result of thb_spec_batch: Array2<f32> (T, B) of calc_spec
 * (spectrogram.rs:187-212).  spec_host, when non-NULL on input, receives the dB values
 * (row-major (n_frames, n_bins), capacity spec_host_cap floats). */
typedef struct thb_spec_out {
    uint64_t n_frames; /* T computed (frame_count) */
    uint64_t total_frames; /* T of the whole file */
    uint32_t n_bins;   /* B: n_fft/2+1 (Linear) or n_mel (Mel) */
    uint32_t hop, win, n_fft;
    float *spec_host;
    uint64_t spec_host_cap;
} thb_spec_out;

/* ---- context ------------------------------------------------------------------------------ */
/* One ctx per process and device (the reference keeps its state in process globals,
 * lib.rs:36-42).  `cuda_stream` may be NULL (library creates its own non-blocking stream) or a
 * cudaStream_t the caller owns (e.g. torch.cuda.current_stream().cuda_stream). */
int thb_ctx_create(int device, void *cuda_stream, thb_ctx **out);
void thb_ctx_destroy(thb_ctx *ctx);
int thb_set_stream(thb_ctx *ctx, void *cuda_stream);
int thb_synchronize(thb_ctx *ctx);
const char *thb_last_error(const thb_ctx *ctx); /* ctx may be NULL: last global error */
int thb_abi_version(void);

/* pinned host memory for PCM / result buffers */
int thb_host_alloc(size_t bytes, void **out);
int thb_host_free(void *p);

/* ---- host parameter arithmetic (no device needed; integer-exact parity) ------------------- */
/* SpecSetting::calc_framing_params (spectrogram.rs:57-98) */
int thb_framing_params(const thb_setting *s, uint32_t sr, uint64_t *hop, uint64_t *win, uint64_t *n_fft);
/* frame count of perform_stft (stft.rs:16-124): 1 + (N + 2*(W/2) - W) / H */
uint64_t thb_n_frames(uint64_t len, uint64_t win, uint64_t hop);
/* B of calc_spec for this (setting, sr): n_fft/2+1 or n_mel (default rule when n_mel == 0) */
int thb_n_bins(const thb_setting *s, uint32_t sr, uint32_t *n_bins);
/* calc_normalized_win(Hann, win, n_fft) (windows.rs:12-38): out[win] */
int thb_hann_window(uint64_t win, uint64_t n_fft, float *out);
/* calc_mel_fb(sr, n_fft, n_mel, 0, None, true) / calc_mel_fb_default (src-common/src/lib.rs:46-103):
 * out is (n_fft/2+1, n_mel) row-major; n_mel == 0 -> default rule; *n_mel_out gets the count.
 * out may be NULL to query n_mel only. */
int thb_mel_fb(uint32_t sr, uint64_t n_fft, uint32_t n_mel, float *out, uint32_t *n_mel_out);
/* Diagnostic: replays the warp schedule the n_fft == 2048 kernels follow for the sparse mel product (the same
 * contraction as spectrogram.rs:207, zeros skipped) on the host and writes the (n_fft/2+1, n_mel) matrix it
 * amounts to -- it must equal thb_mel_fb's.  stats[0..3] = {schedule valid, groups, steps per frame, shared-memory
 * wavefronts lost to bank conflicts per frame pair}.  Returns THB_ERR_UNSUPPORTED when there is no schedule. */
int thb_mel_schedule_replay(uint32_t sr, uint64_t n_fft, uint32_t n_mel, float *out, uint32_t stats[4]);
/* FreqScale::hz_range_to_idx (src-common/src/lib.rs:144-159) */
int thb_hz_range_to_idx(uint32_t freq_scale, float hz0, float hz1, uint32_t sr, uint64_t n_bins,
                        uint64_t *i0, uint64_t *i1);

/* ---- TrackManager::update_specs (src-tauri/src/core/mod.rs:137-164) -------------------------
 * = SpectrogramAnalyzer::prepare (spectrogram.rs:116-154) + calc_spec per (id, ch)
 * (spectrogram.rs:187-212 -> perform_stft stft.rs:16-124 -> norm -> [mel dot] -> dB
 * decibel.rs:170-214).  The dB spectrogram of every track stays resident on the device under
 * (id, ch), replacing any previous one, together with its local (min, max). */
int thb_spec_batch(thb_ctx *ctx, const thb_track *tracks, size_t n, const thb_setting *setting,
                   thb_spec_out *outs /* [n], may be NULL */);
/* Install a caller-computed dB spectrogram (T, B) row-major f32 (host or device memory) under
 * (id, ch) as if thb_spec_batch had produced it: its (min, max) is reduced on the device and it
 * takes part in thb_update_spec_imgs.  Lets a host restore cached specs, and lets tests drive the
 * image path with the reference's own vectors (drawing.rs:43-56). */
int thb_spec_put(thb_ctx *ctx, uint64_t id, uint32_t ch, uint32_t sr, uint32_t freq_scale,
                 const float *spec, uint64_t n_frames, uint32_t n_bins);
/* copy a retained dB spectrogram to the host: out (T, B) row-major, cap in floats */
int thb_spec_read(thb_ctx *ctx, uint64_t id, uint32_t ch, float *out, uint64_t cap,
                  uint64_t *n_frames, uint32_t *n_bins);
/* device pointer of a retained dB spectrogram (T, B) row-major f32 (valid until release/replace) */
int thb_spec_device_ptr(thb_ctx *ctx, uint64_t id, uint32_t ch, const float **dptr,
                        uint64_t *n_frames, uint32_t *n_bins);
/* local (min, max) of one retained spectrogram = find_min_max(spec) (simd.rs:14-36) */
int thb_spec_minmax(thb_ctx *ctx, uint64_t id, uint32_t ch, float *mn, float *mx);
/* TrackManager::remove_tracks (mod.rs:86-100) */
int thb_release(thb_ctx *ctx, uint64_t id, uint32_t ch);
int thb_release_all(thb_ctx *ctx);

/* SpectrogramAnalyzer::prepare / retain (spectrogram.rs:116-185): the per-(sr, win, n_fft[, mel]) plan cache (window,
 * twiddles, mel schedules; KBs each, built on first use by thb_spec_batch).  thb_plans_prepare builds the plans of
 * `setting` for the n sample rates ahead of time; thb_plans_retain drops every cached plan that `setting` does not need
 * for those sample rates -- what TrackManager does after remove_tracks / set_setting (mod.rs:96-99,110-112) -- and
 * reports how many are left. */
int thb_plans_prepare(thb_ctx *ctx, const thb_setting *setting, const uint32_t *srs, size_t n);
/* Diagnostic: which STFT kernel family thb_spec_batch runs for (setting, sr) -- a plan whose tables outgrow a fast
 * kernel's shared memory silently takes a slower one, and the tests pin the expected family of every BASELINE
 * configuration.  *family: THB_KERNEL_GENERIC .. THB_KERNEL_WARP; *mel_schedule: 0 linear, 1 bin-major, 2 band-major. */
#define THB_KERNEL_GENERIC 0u  /* thb_stft_generic.cu */
#define THB_KERNEL_FAST 1u     /* n_fft 2048, scalar warp-per-frame only */
#define THB_KERNEL_PAIR 2u     /* n_fft 2048, packed frame pairs + scalar twin */
#define THB_KERNEL_BIG 3u      /* n_fft 1024 / 4096 / 8192 / 16384, one frame pair per CTA */
#define THB_KERNEL_WARP 4u     /* n_fft 1024 / 512, packed frame pairs + scalar twin */
int thb_plan_kernel(thb_ctx *ctx, const thb_setting *setting, uint32_t sr, uint32_t *family, uint32_t *mel_schedule);
int thb_plans_retain(thb_ctx *ctx, const thb_setting *setting, const uint32_t *srs, size_t n, size_t *n_left);

/* ---- TrackManager::update_spec_imgs (mod.rs:168-230) ----------------------------------------
 * thb_minmax_global: global (min, max) over every retained spectrogram (mod.rs:169-178), reduced
 * across ranks with one exchange of {max, -min} when a communicator is attached (one kernel over the peers'
 * NVLink-mapped memory on one box, see thb_comm_init; ncclAllReduce(max) otherwise), then
 * max <- min(max, 0); min <- max(min, max - dB_range) (mod.rs:179-180). */
int thb_minmax_global(thb_ctx *ctx, float dB_range, float *min_dB, float *max_dB);
/* convert_spectrogram_to_img (visualize/drawing.rs:4-33) for one retained spectrogram:
 * out is (i1 - i0, T) row-major u16 on the HOST. */
int thb_spec_to_img(thb_ctx *ctx, uint64_t id, uint32_t ch, uint64_t i0, uint64_t i1, float min_dB,
                    float max_dB, uint32_t colormap_length, uint16_t *out, uint64_t cap);
/* The whole of update_spec_imgs: global min/max over ALL retained spectrograms (+ all-reduce),
 * i_freq_range from hz_range_to_idx((0, max_sr/2), sr, B) (mod.rs:208-213; max_sr == 0 -> max over
 * retained tracks), quantise + transpose into device-resident images.  `only_ids` (n_only track
 * ids) restricts the quantise step to those tracks -- the reference's `ids_need_update` when
 * neither the dB range nor max_sr moved (mod.rs:194-203); NULL = every track. */
int thb_update_spec_imgs(thb_ctx *ctx, float dB_range, uint32_t colormap_length, uint32_t max_sr,
                         const uint64_t *only_ids, size_t n_only, float *min_dB, float *max_dB);
/* With min_dB == max_dB == NULL thb_update_spec_imgs only QUEUES its work (reduce, all-reduce, quantise) on the ctx
 * stream and returns: a host that re-analyses in a loop keeps the device busy back to back.  thb_range_get waits for
 * the stream and returns the (min_dB, max_dB) of the last update.
 * MULTI-GPU CONTRACT: thb_update_spec_imgs and thb_minmax_global issue exactly ONE collective each, whatever
 * `only_ids` holds; every rank of the communicator must make the same sequence of these two calls.  When only the
 * quantise step is needed for an already reduced range (the reference's `ids_need_update` branch with an unchanged
 * range, mod.rs:194-203, or set_colormap_length, mod.rs:123-131), thb_update_spec_imgs_range does it with the
 * caller's (min_dB, max_dB) and NO collective, so ranks with different track lists may call it independently. */
int thb_range_get(thb_ctx *ctx, float *min_dB, float *max_dB);
int thb_update_spec_imgs_range(thb_ctx *ctx, float min_dB, float max_dB, uint32_t colormap_length, uint32_t max_sr,
                               const uint64_t *only_ids, size_t n_only);
/* TrackManager::get_spectrogram (mod.rs:133-135): copy image (H, W = T) to the host */
int thb_img_read(thb_ctx *ctx, uint64_t id, uint32_t ch, uint16_t *out, uint64_t cap,
                 uint64_t *height, uint64_t *width);
/* the images of n (id, ch) in one go: every copy is queued before the one wait (what a redraw of all tracks asks
 * for; TrackManager.spec_imgs as a whole).  outs[i] is HOST memory for (H_i, T_i) u16, caps[i] its size in pixels. */
int thb_img_read_batch(thb_ctx *ctx, size_t n, const uint64_t *ids, const uint32_t *chs, uint16_t *const *outs,
                       const uint64_t *caps);
/* Install a caller-provided image (height, width = T) u16 row-major (host or device memory) as the retained image of
 * (id, ch), whose spectrogram must exist with T == width: lets a host restore cached images, and lets tests drive the
 * tile path with the reference's own vectors (render_tiles.rs:435-471). */
int thb_img_put(thb_ctx *ctx, uint64_t id, uint32_t ch, const uint16_t *img, uint64_t height, uint64_t width);
/* device view of a retained image: rows are `pitch` u16 apart */
int thb_img_device_ptr(thb_ctx *ctx, uint64_t id, uint32_t ch, const uint16_t **dptr,
                       uint64_t *height, uint64_t *width, uint64_t *pitch);

/* ---- encode_spectrogram_tile (src-tauri/src/core/render_tiles.rs:281-393) ---------------
 * What RenderTileCache::spectrogram_tile (render_tiles.rs:170-188, called by get_spectrogram_tile, lib.rs:369-389)
 * returns for the retained image of (id, ch): 40-byte header {u64 revision, u32 width, height, level_x, level_y,
 * tile_x, tile_y, origin_x, origin_y} (little endian), then width x height RGBA pixels, the LAST row of the tile
 * first (high frequencies first).  The tile is the 512 x 512 core at (tile_x, tile_y) of the level of detail
 * (ceil(W / 2^level_x), ceil(H / 2^level_y)) plus a gutter of 4 pixels, resampled from the u16 image with the
 * separable Lanczos3 convolution of fast_image_resize 6.0.0 (U16 pixels; third-party arithmetic restated from the
 * crate's published algorithm) and mapped through `colormap_rgba` (RGBA bytes, >= 1 colour):
 * index = (value * (colours - 1) + 32767) / 65535.  `out` is HOST memory; out == NULL (cap 0) queries the size.
 * thb_spectrogram_tile_geometry: geo = {lod_width, lod_height, origin_x, origin_y, width, height}
 * (render_tiles.rs:290-312; host arithmetic, no device needed). */
int thb_spectrogram_tile_geometry(uint64_t height, uint64_t width, uint32_t level_x, uint32_t level_y, uint32_t tile_x,
                                  uint32_t tile_y, uint64_t geo[6]);
int thb_spectrogram_tile(thb_ctx *ctx, uint64_t id, uint32_t ch, const uint8_t *colormap_rgba, size_t colormap_bytes,
                         uint64_t revision, uint32_t level_x, uint32_t level_y, uint32_t tile_x, uint32_t tile_y, uint8_t *out,
                         size_t cap, size_t *written);
/* n tiles (any mix of tracks, levels and positions) in one pair of launches: what a redraw of the viewport asks for */
typedef struct thb_spec_tile_req {
    uint64_t id;
    uint32_t ch, level_x, level_y, tile_x, tile_y;
    uint32_t reserved;
    uint8_t *out;    /* HOST memory, or NULL to query `written` */
    size_t cap;
    size_t written;  /* out: 40 + width * height * 4 */
} thb_spec_tile_req;
int thb_spectrogram_tile_batch(thb_ctx *ctx, const uint8_t *colormap_rgba, size_t colormap_bytes, uint64_t revision,
                               thb_spec_tile_req *reqs, size_t n);

/* ---- encode_waveform_tile (src-tauri/src/core/render_tiles.rs:232-279) ----------------------
 * Byte-identical wire format: u64 revision, u32 bin_count, u32 samples_per_bin, u32 tile_index,
 * u32 0, then per bin f32 min, f32 max, f32 mean (little endian).  `out` is HOST memory. */
int thb_waveform_tile(thb_ctx *ctx, const float *pcm, uint64_t len, uint64_t revision, uint32_t level,
                      uint32_t tile_index, uint8_t *out, size_t cap, size_t *written);
/* Host PCM handed to thb_waveform_tile stays on the device under the key (pcm pointer, len, revision) -- the key of the
 * reference's own tile cache (render_tiles.rs:124-169) -- filled granule by granule with exactly the samples tile calls
 * touch, so a later tile of the same channel does not cross PCIe again.  A changed channel must come with a new
 * revision (as in the reference) or a new pointer.  Bounded by the environment variable THB_PCM_CACHE_MB (default
 * 4096; 0 turns the cache off), least recently used first. */
int thb_pcm_cache_stats(thb_ctx *ctx, uint64_t *entries, uint64_t *bytes, uint64_t *hits, uint64_t *misses);
int thb_pcm_cache_clear(thb_ctx *ctx);
/* every tile of one level, concatenated in tile order (what a full redraw at that zoom asks for) */
int thb_waveform_level(thb_ctx *ctx, const float *pcm, uint64_t len, uint64_t revision, uint32_t level,
                       uint8_t *out, size_t cap, size_t *written);
/* same, for n channels in one launch; the encoded levels stay on the device (dev_out[i], valid until
 * the next call) and are copied to host_out[i] when host_out != NULL */
int thb_waveform_level_batch(thb_ctx *ctx, const thb_track *tracks, size_t n, uint64_t revision,
                             uint32_t level, uint8_t **host_out, const size_t *caps, size_t *written,
                             const uint8_t **dev_out);
uint64_t thb_waveform_level_bytes(uint64_t len, uint32_t level);

/* ---- StatCalculator::calc, level part (src-tauri/src/core/dynamics/stats.rs:56-85) ----
 * thb_channel_stats: for each of n channels (thb_track: pcm, len, pcm_format; host or device) the two reductions the
 * reference runs over the PCM: sum_squares (simd.rs:113-134,820-832) and abs_max (simd.rs:161-183,935-937).
 * thb_audio_stats: the scalar arithmetic on top for ONE track of n_ch channels: mean_squared = sum of the channels'
 * sums / n_elem, rms_dB = 10 log10 (0 -> -inf), max_peak, max_peak_dB = 20 log10 (host code, no device needed).
 * The loudness leg (EBU R128, a sequential filter) is not part of this library. */
typedef struct thb_audio_stats_t {
    float mean_squared;
    float rms_dB;
    float max_peak;
    float max_peak_dB;
} thb_audio_stats_t;
int thb_channel_stats(thb_ctx *ctx, const thb_track *channels, size_t n, float *sum_squares /* [n] */,
                      float *abs_max /* [n] */);
int thb_audio_stats(const float *sum_squares, const float *abs_max, const uint64_t *lens, size_t n_ch,
                    thb_audio_stats_t *out);

/* ---- gain normalisation + guard clipping ahead of the analysis --------------------------
 * thb_normalize_gain: Normalize::normalize_default's gain for a target (dynamics/normalize.rs:23-45), from the ORIGINAL
 * audio's statistics (host arithmetic, f32 like the reference; global_lufs is the reference's f64, narrowed first).
 * thb_apply_gain: AudioTrack::apply_gain (track.rs:152-171: y = gain * original) + the guard clipping of Audio::mutate
 * (audio.rs:49-63) + the level statistics `mutate` recomputes, for n channels in one pass over the samples.  Channels
 * with the same `id` form one Audio (they must carry the same gain): ReduceGlobalLevel takes its peak over all of
 * them (audio.rs:146-160).  A non-finite or unit gain restores the original (track.rs:160-161).  `pcm`, `out` and
 * `before_clip` may each be host or device memory; `out` may equal `pcm` (f32).  `before_clip` (optional) receives
 * GuardClippingResult::WavBeforeClip in Clip mode (what channel_for_drawing shows, audio.rs:71-79).  The limiter mode
 * (audio.rs:161-178, limiter.rs:47-176) is a sequential recurrence and is not provided: THB_ERR_UNSUPPORTED. */
#define THB_NORM_OFF 0u
#define THB_NORM_LUFS 1u
#define THB_NORM_RMS_DB 2u
#define THB_NORM_PEAK_DB 3u
float thb_normalize_gain(uint32_t target_kind, float target, double global_lufs, float rms_dB, float max_peak_dB);

#define THB_GUARD_CLIP 0u                 /* GuardClippingMode::Clip (dynamics/guardclipping.rs:7-12) */
#define THB_GUARD_REDUCE_GLOBAL_LEVEL 1u
#define THB_GUARD_LIMITER 2u              /* not provided */
typedef struct thb_gain_channel {
    const void *pcm;     /* the ORIGINAL samples: f32, or i16 when pcm_format == THB_PCM_I16 */
    uint64_t len;
    uint64_t id;         /* track id: groups the channels of one Audio */
    uint32_t pcm_format;
    float gain;
    float *out;          /* len floats: Audio.wavs after gain + guard clipping */
    float *before_clip;  /* NULL, or len floats */
} thb_gain_channel;
typedef struct thb_gain_result {
    float global_gain;           /* GuardClippingResult::GlobalGain (1 unless ReduceGlobalLevel reduced) */
    float max_reduction_gain_dB; /* GuardClippingStats (dynamics/stats.rs:110-158) of this channel */
    uint64_t reduction_cnt;
    float sum_squares;           /* of the output channel: the inputs of thb_audio_stats */
    float abs_max;
} thb_gain_result;
int thb_apply_gain(thb_ctx *ctx, const thb_gain_channel *channels, size_t n, uint32_t mode, thb_gain_result *results /* [n] */);

/* ---- multi-GPU: one process per GPU, one communicator per box ---------
 * NCCL is dlopen'ed ("libnccl.so.2") on first use.  Rank 0 creates the id, the host program
 * distributes the 128 bytes (torch.distributed / MPI / a file), every rank calls thb_comm_init. */
int thb_comm_unique_id(uint8_t id[128]);
int thb_comm_init(thb_ctx *ctx, int n_ranks, int rank, const uint8_t id[128]);
int thb_comm_destroy(thb_ctx *ctx);
/* thb_comm_init also maps, through CUDA IPC, a 512-byte exchange buffer of every rank into every other rank (the handles
 * travel over the new communicator): the global dB range then takes ONE kernel that stores {max, -min} into the peers'
 * buffers over NVLink and polls its own, instead of reduce + ncclAllReduce + finalize.  If any mapping fails on any rank
 * (no peer access, IPC not permitted, more than 16 ranks, or THB_PEER_EXCHANGE=0) every rank keeps the NCCL path.
 * Returns 1 when the peer-memory exchange is in use. */
int thb_comm_peer_exchange(const thb_ctx *ctx);

/* ---- measurement support --------------------------------------------------------------------
 * Per-kernel CUDA-event timing on the ctx stream and a launch counter (bench.py's roofline /
 * gpu_launches).  Kernel names: "stft_mel_db", "stft_lin_db" (the main STFT kernel of a batch),
 * "stft_mel_db_edges", "stft_lin_db_edges" (file-edge frames and rescued tiles, when the main kernel
 * leaves them to the scalar one), "minmax_reduce", "minmax_array", "spec_to_img", "envelope", "channel_stats", "gain_peak", "gain_apply", "spectrogram_tile". */
int thb_profile_enable(thb_ctx *ctx, int on);
int thb_profile_reset(thb_ctx *ctx);
int thb_profile_get(thb_ctx *ctx, const char *kernel, double *total_ms, uint64_t *launches);
uint64_t thb_launch_count(const thb_ctx *ctx);
/* deterministic integer-arithmetic synthetic PCM, written to DEVICE memory
 * `dev_out` (len floats); thesia_b200/synth.py restates it in numpy bit for bit. */
int thb_synth_pcm(thb_ctx *ctx, float *dev_out, uint64_t len, uint32_t sr, uint32_t track, uint32_t channel,
                  uint32_t flags);
#define THB_SYNTH_LOUD 1u     /* scale x32 (f32 audio above 0 dBFS) so that max_dB hits the min(max, 0) clamp */
#define THB_SYNTH_ZERO_GAP 2u /* one second of exact zeros (produces -inf frames) */

#ifdef __cplusplus
}
#endif
#endif /* THESIA_B200_H */
