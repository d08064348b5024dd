// thesia_host.hpp -- the host side of the analysis path in C++17, above the C ABI (include/thesia_b200.h).
//
// The reference's host code is Rust (src-tauri/src/core); there is no Rust toolchain in this build, so the types a
// maintainer would keep on the Rust side of the FFI are mirrored here with the reference's names, argument meaning
// and update rules:
//
//   SpecSetting, FreqScale     src-tauri/src/core/spectrogram.rs:30-98, src-common/src/lib.rs:106-159
//   TrackList (analysis part)  src-tauri/src/core/track.rs:199-437
//   TrackManager               src-tauri/src/core/mod.rs:33-231
//   encode_waveform_tile       src-tauri/src/core/render_tiles.rs:232-279
//
// Nothing here touches a sample: every call ends in libthesia_b200.so (CUDA kernels).  The reference's analysis
// path has no error channel (it unwraps / panics, stft.rs:47); the mirror throws thb::host::Error instead.
// Header-only; link with -lthesia_b200.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/thesia_b200.h"

namespace thb {
namespace host {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc, const thb_ctx *ctx = nullptr) {
    if (rc != THB_OK) {
        const char *msg = thb_last_error(ctx);
        throw Error(rc, std::string("thesia_b200 error ") + std::to_string(rc) + ": " + (msg ? msg : ""));
    }
}

// FreqScale (src-common/src/lib.rs:106-110)
enum class FreqScale : uint32_t { Linear = THB_FREQ_LINEAR, Mel = THB_FREQ_MEL };

// FreqScale::hz_range_to_idx (src-common/src/lib.rs:144-159)
inline std::pair<uint64_t, uint64_t> hz_range_to_idx(FreqScale scale, std::pair<float, float> hz_range, uint32_t sr,
                                                     uint64_t n_freqs) {
    uint64_t i0 = 0, i1 = 0;
    check(thb_hz_range_to_idx(static_cast<uint32_t>(scale), hz_range.first, hz_range.second, sr, n_freqs, &i0, &i1));
    return {i0, i1};
}

// SrWinNfft (spectrogram.rs:40-45)
struct SrWinNfft {
    uint32_t sr = 0;
    uint64_t win_length = 0, n_fft = 0;
    bool operator<(const SrWinNfft &o) const { return std::tie(sr, win_length, n_fft) < std::tie(o.sr, o.win_length, o.n_fft); }
};

// SpecSetting (spectrogram.rs:30-38; Default at :47-54)
struct SpecSetting {
    double win_ms = 40.0;
    uint32_t t_overlap = 4;
    uint32_t f_overlap = 1;
    FreqScale freq_scale = FreqScale::Mel;
    uint32_t n_mel = 0;  // 0 = calc_mel_fb_default's rule (what TrackManager always uses)

    thb_setting c() const { return thb_setting{win_ms, t_overlap, f_overlap, static_cast<uint32_t>(freq_scale), n_mel}; }

    // (hop, win, n_fft) -- spectrogram.rs:57-98
    std::tuple<uint64_t, uint64_t, uint64_t> calc_framing_params(uint32_t sr) const {
        const thb_setting s = c();
        uint64_t hop = 0, win = 0, n_fft = 0;
        check(thb_framing_params(&s, sr, &hop, &win, &n_fft));
        return {hop, win, n_fft};
    }
    uint64_t calc_hop_length(uint32_t sr) const { return std::get<0>(calc_framing_params(sr)); }
    uint64_t calc_win_length(uint32_t sr) const { return std::get<1>(calc_framing_params(sr)); }
    uint64_t calc_n_fft(uint32_t sr) const { return std::get<2>(calc_framing_params(sr)); }
    SrWinNfft calc_sr_win_nfft(uint32_t sr) const {
        const auto [hop, win, n_fft] = calc_framing_params(sr);
        (void)hop;
        return SrWinNfft{sr, win, n_fft};
    }
    uint32_t n_bins(uint32_t sr) const {
        const thb_setting s = c();
        uint32_t b = 0;
        check(thb_n_bins(&s, sr, &b));
        return b;
    }
};

inline uint64_t n_frames(uint64_t len, uint64_t win, uint64_t hop) { return thb_n_frames(len, win, hop); }

// calc_normalized_win(Hann, win, n_fft) (windows.rs:12-38)
inline std::vector<float> calc_normalized_win(uint64_t win, uint64_t n_fft) {
    std::vector<float> w(win);
    check(thb_hann_window(win, n_fft, w.data()));
    return w;
}

// calc_mel_fb / calc_mel_fb_default (src-common/src/lib.rs:46-103): (n_fft/2+1, n_mel) row-major
struct MelFb {
    uint64_t n_freq = 0;
    uint32_t n_mel = 0;
    std::vector<float> w;
    float at(uint64_t k, uint32_t m) const { return w[k * n_mel + m]; }
};
inline MelFb calc_mel_fb(uint32_t sr, uint64_t n_fft, uint32_t n_mel /* 0 = default rule */) {
    MelFb fb;
    fb.n_freq = n_fft / 2 + 1;
    check(thb_mel_fb(sr, n_fft, n_mel, nullptr, &fb.n_mel));
    fb.w.assign(fb.n_freq * fb.n_mel, 0.0f);
    check(thb_mel_fb(sr, n_fft, n_mel, fb.w.data(), &fb.n_mel));
    return fb;
}
inline MelFb calc_mel_fb_default(uint32_t sr, uint64_t n_fft) { return calc_mel_fb(sr, n_fft, 0); }

using IdCh = std::pair<uint64_t, uint32_t>;

// Owner of the device context (the reference keeps its state in process globals, lib.rs:36-42).
class Context {
public:
    explicit Context(int device = 0, void *cuda_stream = nullptr) { check(thb_ctx_create(device, cuda_stream, &h_)); }
    ~Context() {
        if (h_) thb_ctx_destroy(h_);
    }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    thb_ctx *get() const { return h_; }

private:
    thb_ctx *h_ = nullptr;
};

// Audio (audio.rs:24-30): `wavs` is (n_ch, len) C-contiguous f32, a channel is one row.
struct Audio {
    std::vector<float> wavs;
    uint32_t n_ch = 0;
    uint64_t len = 0;
    uint32_t sr = 0;
    const float *channel(uint32_t ch) const { return wavs.data() + static_cast<size_t>(ch) * len; }
};

// The part of TrackList (track.rs:199-437) the analysis path reads.
class TrackList {
public:
    std::vector<uint64_t> add_tracks(const std::vector<uint64_t> &id_list, std::vector<Audio> audios) {
        for (size_t i = 0; i < id_list.size(); i++) tracks_[id_list[i]] = std::move(audios[i]);
        return id_list;
    }
    std::vector<IdCh> remove_tracks(const std::vector<uint64_t> &id_list) {
        std::vector<IdCh> removed;
        for (uint64_t id : id_list) {
            auto it = tracks_.find(id);
            if (it == tracks_.end()) continue;
            for (uint32_t ch = 0; ch < it->second.n_ch; ch++) removed.emplace_back(id, ch);
            tracks_.erase(it);
        }
        return removed;
    }
    bool has(uint64_t id) const { return tracks_.count(id) != 0; }
    const Audio &get(uint64_t id) const { return tracks_.at(id); }
    std::vector<uint64_t> all_ids() const {
        std::vector<uint64_t> v;
        for (auto &kv : tracks_) v.push_back(kv.first);
        return v;
    }
    std::set<uint64_t> all_id_set() const {
        std::set<uint64_t> s;
        for (auto &kv : tracks_) s.insert(kv.first);
        return s;
    }
    uint32_t max_sr() const {
        uint32_t m = 0;
        for (auto &kv : tracks_) m = std::max(m, kv.second.sr);
        return m;
    }
    std::vector<IdCh> id_ch_tuples_from(const std::vector<uint64_t> &id_list) const {
        std::vector<IdCh> v;
        for (uint64_t id : id_list) {
            auto it = tracks_.find(id);
            if (it == tracks_.end()) continue;
            for (uint32_t ch = 0; ch < it->second.n_ch; ch++) v.emplace_back(id, ch);
        }
        return v;
    }
    std::vector<IdCh> id_ch_tuples() const { return id_ch_tuples_from(all_ids()); }
    // construct_sr_win_nfft_set (track.rs): the analyzer plans the batch will need
    std::set<SrWinNfft> construct_sr_win_nfft_set(const std::vector<uint64_t> &ids, const SpecSetting &setting) const {
        std::set<SrWinNfft> s;
        for (uint64_t id : ids)
            if (has(id)) s.insert(setting.calc_sr_win_nfft(get(id).sr));
        return s;
    }

private:
    std::map<uint64_t, Audio> tracks_;
};

// Array2<u16> (H, W) of convert_spectrogram_to_img (drawing.rs:4-33)
struct SpecImg {
    uint64_t height = 0, width = 0;
    std::vector<uint16_t> px;
    uint16_t at(uint64_t row, uint64_t col) const { return px[row * width + col]; }
};
// Array2<f32> (T, B) of calc_spec (spectrogram.rs:187-212)
struct Spec {
    uint64_t n_frames = 0;
    uint32_t n_bins = 0;
    std::vector<float> dB;
    float at(uint64_t t, uint32_t b) const { return dB[t * n_bins + b]; }
};

// TrackManager (mod.rs:33-231).  `specs` and `spec_imgs` are resident on the device; get_spectrogram copies one out.
class TrackManager {
public:
    float max_dB = -std::numeric_limits<float>::infinity();
    float min_dB = std::numeric_limits<float>::infinity();
    uint32_t max_sr = 0;
    SpecSetting setting;
    float dB_range = 100.0f;
    uint32_t colormap_length = 258;

    explicit TrackManager(Context &ctx) : ctx_(ctx) {}

    // mod.rs:62-84
    void add_tracks(const TrackList &tracklist, const std::vector<uint64_t> &added_ids) {
        update_specs(tracklist, tracklist.id_ch_tuples_from(added_ids));
        no_spec_img_ids_.insert(no_spec_img_ids_.end(), added_ids.begin(), added_ids.end());
    }
    void reload_tracks(const TrackList &tracklist, const std::vector<uint64_t> &reloaded_ids) {
        update_specs(tracklist, tracklist.id_ch_tuples_from(reloaded_ids));
        no_spec_img_ids_.insert(no_spec_img_ids_.end(), reloaded_ids.begin(), reloaded_ids.end());
    }
    // mod.rs:86-100
    // SpectrogramAnalyzer::retain (spectrogram.rs:156-185) for the sample rates still in the track list
    size_t retain_plans(const TrackList &tracklist) {
        std::set<uint32_t> srs;
        for (uint64_t id : tracklist.all_ids()) srs.insert(tracklist.get(id).sr);
        const std::vector<uint32_t> v(srs.begin(), srs.end());
        const thb_setting s = setting.c();
        size_t left = 0;
        check(thb_plans_retain(ctx_.get(), &s, v.data(), v.size(), &left), ctx_.get());
        return left;
    }
    void remove_tracks(const TrackList &tracklist, const std::vector<IdCh> &removed_id_ch_tuples) {
        for (const IdCh &tup : removed_id_ch_tuples) {
            if (!specs_.count(tup)) continue;
            check(thb_release(ctx_.get(), tup.first, tup.second), ctx_.get());
            specs_.erase(tup);
            spec_imgs_.erase(tup);
        }
        retain_plans(tracklist);  // spec_analyzer.retain(construct_all_sr_win_nfft_set(setting), freq_scale), mod.rs:96-99
    }
    // mod.rs:102-105
    std::pair<std::set<uint64_t>, uint32_t> apply_track_list_changes(const TrackList &tracklist) {
        std::set<uint64_t> s = update_spec_imgs(tracklist, false);
        return {std::move(s), max_sr};
    }
    // mod.rs:107-121
    void set_setting(const TrackList &tracklist, const SpecSetting &new_setting) {
        setting = new_setting;
        retain_plans(tracklist);  // mod.rs:110-112
        update_specs(tracklist, tracklist.id_ch_tuples());
        update_spec_imgs(tracklist, true);
    }
    void update_all_specs_imgs(const TrackList &tracklist) {
        update_specs(tracklist, tracklist.id_ch_tuples());
        update_spec_imgs(tracklist, true);
    }
    // mod.rs:123-131: only the quantise step runs again, the dB spectrograms stay resident
    void set_dB_range(const TrackList &tracklist, float new_dB_range) {
        dB_range = new_dB_range;
        update_spec_imgs(tracklist, true);
    }
    void set_colormap_length(const TrackList &tracklist, uint32_t new_colormap_length) {
        colormap_length = new_colormap_length;
        update_spec_imgs(tracklist, true);
    }
    // mod.rs:133-135
    std::optional<SpecImg> get_spectrogram(IdCh id_ch) const {
        if (!spec_imgs_.count(id_ch)) return std::nullopt;
        SpecImg img;
        check(thb_img_read(ctx_.get(), id_ch.first, id_ch.second, nullptr, 0, &img.height, &img.width), ctx_.get());
        img.px.assign(img.height * img.width, 0);
        check(thb_img_read(ctx_.get(), id_ch.first, id_ch.second, img.px.data(), img.px.size(), nullptr, nullptr), ctx_.get());
        return img;
    }
    // the reference's private `specs` map, exposed for parity checks
    std::optional<Spec> get_spec(IdCh id_ch) const {
        if (!specs_.count(id_ch)) return std::nullopt;
        Spec sp;
        check(thb_spec_read(ctx_.get(), id_ch.first, id_ch.second, nullptr, 0, &sp.n_frames, &sp.n_bins), ctx_.get());
        sp.dB.assign(sp.n_frames * sp.n_bins, 0.0f);
        check(thb_spec_read(ctx_.get(), id_ch.first, id_ch.second, sp.dB.data(), sp.dB.size(), nullptr, nullptr), ctx_.get());
        return sp;
    }
    bool has_spec_img(IdCh id_ch) const { return spec_imgs_.count(id_ch) != 0; }

private:
    // mod.rs:137-164: one batch over every (id, ch)
    void update_specs(const TrackList &tracklist, const std::vector<IdCh> &id_ch_tuples) {
        std::vector<thb_track> tracks;
        tracks.reserve(id_ch_tuples.size());
        for (const IdCh &t : id_ch_tuples) {
            const Audio &a = tracklist.get(t.first);
            thb_track tr{};
            tr.pcm = a.channel(t.second);
            tr.len = a.len;
            tr.id = t.first;
            tr.ch = t.second;
            tr.sr = a.sr;
            tracks.push_back(tr);
        }
        const thb_setting s = setting.c();
        check(thb_spec_batch(ctx_.get(), tracks.data(), tracks.size(), &s, nullptr), ctx_.get());
        for (const IdCh &t : id_ch_tuples) specs_.insert(t);
    }
    // mod.rs:168-230
    std::set<uint64_t> update_spec_imgs(const TrackList &tracklist, bool force_update_all) {
        // the ONE collective of an update (when a communicator is attached): every rank makes it, whatever its own track
        // list needs afterwards; the quantise step below uses the reduced range and no collective
        float mn = 0.0f, mx = 0.0f;
        check(thb_minmax_global(ctx_.get(), dB_range, &mn, &mx), ctx_.get());
        bool need_update_all = force_update_all;
        if (max_dB != mx) {
            max_dB = mx;
            need_update_all = true;
        }
        if (min_dB != mn) {
            min_dB = mn;
            need_update_all = true;
        }
        const uint32_t new_max_sr = tracklist.max_sr();
        if (max_sr != new_max_sr) {
            max_sr = new_max_sr;
            need_update_all = true;
        }
        std::set<uint64_t> ids_need_update;
        if (need_update_all)
            ids_need_update = tracklist.all_id_set();
        else
            ids_need_update.insert(no_spec_img_ids_.begin(), no_spec_img_ids_.end());
        no_spec_img_ids_.clear();
        if (!ids_need_update.empty()) {
            if (need_update_all) spec_imgs_.clear();
            const std::vector<uint64_t> only(ids_need_update.begin(), ids_need_update.end());
            check(thb_update_spec_imgs_range(ctx_.get(), min_dB, max_dB, colormap_length, max_sr, need_update_all ? nullptr : only.data(),
                                             need_update_all ? 0 : only.size()),
                  ctx_.get());
            for (const IdCh &k : specs_)
                if (ids_need_update.count(k.first)) spec_imgs_.insert(k);
        }
        return ids_need_update;
    }

    Context &ctx_;
    std::set<IdCh> specs_, spec_imgs_;
    std::vector<uint64_t> no_spec_img_ids_;
};

// AudioStats of StatCalculator::calc (dynamics/stats.rs:56-85) without the loudness leg
struct AudioStats {
    float rms_dB = 0.0f, max_peak = 0.0f, max_peak_dB = 0.0f;
};
inline AudioStats calc_stats(Context &ctx, const Audio &audio) {
    std::vector<thb_track> chans(audio.n_ch);
    std::vector<uint64_t> lens(audio.n_ch, audio.len);
    for (uint32_t ch = 0; ch < audio.n_ch; ch++) {
        chans[ch] = thb_track{};
        chans[ch].pcm = audio.channel(ch);
        chans[ch].len = audio.len;
        chans[ch].sr = audio.sr;
    }
    std::vector<float> ss(audio.n_ch), mx(audio.n_ch);
    check(thb_channel_stats(ctx.get(), chans.data(), chans.size(), ss.data(), mx.data()), ctx.get());
    thb_audio_stats_t st{};
    check(thb_audio_stats(ss.data(), mx.data(), lens.data(), audio.n_ch, &st));
    return AudioStats{st.rms_dB, st.max_peak, st.max_peak_dB};
}

// ---- spectrogram tiles --------------------------------------------------------------------
// The spectrogram half of RenderTileCache (render_tiles.rs:51-188): colormap + revision, and spectrogram_tile() for
// the image TrackManager retains for (id, ch) (get_spectrogram_tile, lib.rs:369-389).
constexpr uint64_t SPECTROGRAM_TILE_SIZE = 512;
class RenderTileCache {
public:
    explicit RenderTileCache(Context &ctx) : ctx_(ctx) {}
    uint64_t spectrogram_revision = 1;
    std::vector<uint8_t> colormap_rgba = {0, 0, 0, 255, 255, 255, 255, 255};
    void set_colormap(std::vector<uint8_t> rgba) {  // render_tiles.rs:80-85: an invalid table is ignored, the revision still moves
        if (rgba.size() >= 4 && rgba.size() % 4 == 0) colormap_rgba = std::move(rgba);
        invalidate_spectrogram();
    }
    void invalidate_spectrogram() {
        spectrogram_revision += 1;
        if (spectrogram_revision == 0) spectrogram_revision = 1;  // wrapping_add(1).max(1)
    }
    std::vector<uint8_t> spectrogram_tile(uint64_t id, uint32_t ch, uint32_t level_x, uint32_t level_y, uint32_t tile_x, uint32_t tile_y) {
        size_t need = 0;
        check(thb_spectrogram_tile(ctx_.get(), id, ch, colormap_rgba.data(), colormap_rgba.size(), spectrogram_revision, level_x, level_y,
                                   tile_x, tile_y, nullptr, 0, &need), ctx_.get());
        std::vector<uint8_t> out(need);
        check(thb_spectrogram_tile(ctx_.get(), id, ch, colormap_rgba.data(), colormap_rgba.size(), spectrogram_revision, level_x, level_y,
                                   tile_x, tile_y, out.data(), out.size(), &need), ctx_.get());
        return out;
    }

private:
    Context &ctx_;
};

// ---- gain normalisation + guard clipping ----------------------------------------------------
// GuardClippingMode (dynamics/guardclipping.rs:6-12), NormalizeTarget (dynamics/normalize.rs:6-15)
enum class GuardClippingMode : uint32_t { Clip = THB_GUARD_CLIP, ReduceGlobalLevel = THB_GUARD_REDUCE_GLOBAL_LEVEL, Limiter = THB_GUARD_LIMITER };
struct NormalizeTarget {
    enum Kind : uint32_t { Off = THB_NORM_OFF, LUFS = THB_NORM_LUFS, RMSdB = THB_NORM_RMS_DB, PeakdB = THB_NORM_PEAK_DB } kind = Off;
    float target = 0.0f;
};
// GuardClippingStats (dynamics/stats.rs:110-158)
struct GuardClippingStats {
    float max_reduction_gain_dB = 0.0f;
    uint64_t reduction_cnt = 0;
};
// AudioTrack (track.rs:28-171), the part Normalize touches: `original` never changes, `audio` = gain * original after
// guard clipping, with the statistics and guard-clip state Audio::mutate leaves behind (audio.rs:49-63).
struct AudioTrack {
    Audio original, audio;
    AudioStats original_stats, stats;
    double original_global_lufs = 0.0;       // the loudness leg is computed on the CPU by the host program
    std::vector<float> wav_before_clip;      // GuardClippingResult::WavBeforeClip (Clip mode, when anything was scaled)
    float global_gain = 1.0f;                // GuardClippingResult::GlobalGain
    std::vector<GuardClippingStats> guard_clip_stats;

    // Normalize::normalize_default (normalize.rs:23-45) + AudioTrack::apply_gain (track.rs:158-171)
    void normalize(Context &ctx, NormalizeTarget target, GuardClippingMode mode) {
        apply_gain(ctx, thb_normalize_gain(target.kind, target.target, original_global_lufs, original_stats.rms_dB, original_stats.max_peak_dB), mode);
    }
    void apply_gain(Context &ctx, float gain, GuardClippingMode mode) {
        audio = original;
        const bool restore = !std::isfinite(gain) || gain == 1.0f;
        const bool clip = mode == GuardClippingMode::Clip && !restore;
        wav_before_clip.assign(clip ? original.wavs.size() : 0, 0.0f);
        std::vector<thb_gain_channel> chans(original.n_ch);
        for (uint32_t ch = 0; ch < original.n_ch; ch++) {
            chans[ch] = thb_gain_channel{};
            chans[ch].pcm = original.channel(ch);
            chans[ch].len = original.len;
            chans[ch].gain = gain;
            chans[ch].out = audio.wavs.data() + static_cast<size_t>(ch) * audio.len;
            chans[ch].before_clip = clip ? wav_before_clip.data() + static_cast<size_t>(ch) * audio.len : nullptr;
        }
        std::vector<thb_gain_result> res(original.n_ch);
        check(thb_apply_gain(ctx.get(), chans.data(), chans.size(), static_cast<uint32_t>(mode), res.data()), ctx.get());
        guard_clip_stats.assign(original.n_ch, GuardClippingStats{});
        std::vector<float> ss(original.n_ch), mx(original.n_ch);
        std::vector<uint64_t> lens(original.n_ch, original.len);
        for (uint32_t ch = 0; ch < original.n_ch; ch++) {
            guard_clip_stats[ch] = GuardClippingStats{res[ch].max_reduction_gain_dB, res[ch].reduction_cnt};
            ss[ch] = res[ch].sum_squares;
            mx[ch] = res[ch].abs_max;
        }
        global_gain = original.n_ch ? res[0].global_gain : 1.0f;
        thb_audio_stats_t st{};
        check(thb_audio_stats(ss.data(), mx.data(), lens.data(), original.n_ch, &st));
        stats = AudioStats{st.rms_dB, st.max_peak, st.max_peak_dB};
    }
};

// encode_waveform_tile(&[f32], u64, u32, u32) -> Vec<u8> (render_tiles.rs:232-279)
inline std::vector<uint8_t> encode_waveform_tile(Context &ctx, const float *wav, uint64_t len, uint64_t revision, uint32_t level,
                                                 uint32_t tile_index) {
    size_t need = 0;
    check(thb_waveform_tile(ctx.get(), wav, len, revision, level, tile_index, nullptr, 0, &need), ctx.get());
    std::vector<uint8_t> out(need);
    check(thb_waveform_tile(ctx.get(), wav, len, revision, level, tile_index, out.data(), out.size(), &need), ctx.get());
    out.resize(need);
    return out;
}

}  // namespace host
}  // namespace thb
