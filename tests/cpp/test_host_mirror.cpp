// test_host_mirror.cpp -- the C++ host mirror (thesia_b200/host/thesia_host.hpp) against the reference's own
// known-answer tests, through the C ABI only.
//
//   test_host_mirror --cpu   host arithmetic (no device): SpecSetting framing (spectrogram.rs:57-98), Hann window
//                            (windows.rs:88-91 hann_window_works), mel bank (src-common/src/lib.rs:168-232 mel_works,
//                            mel_default_works), hz_range_to_idx, and the loud failure without a device
//   test_host_mirror --gpu   TrackManager flow on a device: stft_works (stft.rs:173-196) through calc_spec's dB,
//                            update rules of mod.rs:168-230, set_dB_range re-quantising only, the image KAT
//                            (drawing.rs:43-56) and the waveform tile KATs (render_tiles.rs:408-433)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <thread>

#include "../../thesia_b200/host/thesia_host.hpp"

using namespace thb::host;

static int g_fail = 0;
#define EXPECT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) {                                                        \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);       \
            g_fail++;                                                         \
        }                                                                     \
    } while (0)

static void test_framing() {
    struct Case { double win_ms; uint32_t sr, t, f; uint64_t hop, win, n_fft; };
    const Case cases[] = {
        {40.0, 48000, 4, 1, 480, 1920, 2048},          // the reference default setting at 48 kHz
        {40.0, 44100, 4, 1, 441, 1764, 2048},
        {2048 / 48.0, 48000, 4, 1, 512, 2048, 2048},   // BASELINE config C1 / C3
        {2048 / 48.0, 48000, 8, 1, 256, 2048, 2048},   // C2
        {16384 / 96.0, 96000, 16, 1, 1024, 16384, 16384},  // C4
        {40.0, 22050, 4, 2, 221, 884, 2048},           // f_overlap doubles n_fft
        {1.0, 8000, 1, 1, 8, 8, 8},
    };
    for (const Case &c : cases) {
        SpecSetting s{c.win_ms, c.t, c.f, FreqScale::Linear, 0};
        const auto [hop, win, n_fft] = s.calc_framing_params(c.sr);
        EXPECT(hop == c.hop && win == c.win && n_fft == c.n_fft);
        EXPECT(s.calc_hop_length(c.sr) == c.hop && s.calc_win_length(c.sr) == c.win);
        const SrWinNfft k = s.calc_sr_win_nfft(c.sr);
        EXPECT(k.sr == c.sr && k.win_length == c.win && k.n_fft == c.n_fft);
        EXPECT(s.n_bins(c.sr) == c.n_fft / 2 + 1);
    }
    // T = 1 + N / H for even windows (stft.rs:50-95)
    EXPECT(n_frames(2113529, 2048, 512) == 4128);
    EXPECT(n_frames(28800000, 2048, 512) == 56251);
    EXPECT(n_frames(2, 8, 8) == 1);
    const SpecSetting d;  // Default (spectrogram.rs:47-54)
    EXPECT(d.win_ms == 40.0 && d.t_overlap == 4 && d.f_overlap == 1 && d.freq_scale == FreqScale::Mel);
}

static void test_window_and_mel() {
    // hann_window_works: hann(4, false) = [0, .5, 1, .5]; calc_normalized_win divides by n_fft
    const std::vector<float> w = calc_normalized_win(4, 4);
    EXPECT(w.size() == 4 && w[0] == 0.0f && w[1] == 0.125f && w[2] == 0.25f && w[3] == 0.125f);
    // mel_works: first filter of calc_mel_fb(24000, 2048, 80)
    const double ans[9] = {0.0, 0.07852016499598029, 0.15704032999196058, 0.23556049498794085, 0.25,
                           0.17147983500401973, 0.09295967000803942, 0.014439505012059144, 0.0};
    const MelFb fb = calc_mel_fb(24000, 2048, 80);
    EXPECT(fb.n_freq == 1025 && fb.n_mel == 80);
    for (int k = 0; k < 9; k++) EXPECT(std::fabs(fb.at(k, 0) - ans[k]) < 5e-6);
    for (uint64_t k = 9; k < fb.n_freq; k++) EXPECT(fb.at(k, 0) == 0.0f);
    // mel_default_works: every default filter is non-empty, one more band is not
    for (uint32_t sr : {8000u, 16000u, 44100u, 48000u, 96000u}) {
        const MelFb d = calc_mel_fb_default(sr, 2048);
        bool all = true;
        for (uint32_t m = 0; m < d.n_mel; m++) {
            float s = 0.0f;
            for (uint64_t k = 0; k < d.n_freq; k++) s += d.at(k, m);
            all = all && s > 0.0f;
        }
        EXPECT(all);
        EXPECT(SpecSetting({2048.0 * 1000.0 / sr, 4, 1, FreqScale::Mel, 0}).n_bins(sr) == d.n_mel);
    }
    EXPECT(calc_mel_fb_default(48000, 2048).n_mel == 347);
    // hz_range_to_idx (lib.rs:144-159)
    EXPECT(hz_range_to_idx(FreqScale::Linear, {0.0f, 24000.0f}, 48000, 1025) == std::make_pair(uint64_t(0), uint64_t(1025)));
    EXPECT(hz_range_to_idx(FreqScale::Linear, {0.0f, 24000.0f}, 24000, 1025).second == 1025 * 2);
    EXPECT(hz_range_to_idx(FreqScale::Mel, {100.0f, 100.0f}, 48000, 128) == std::make_pair(uint64_t(0), uint64_t(0)));
}

static void test_no_device_is_loud() {
    // only meaningful on a box without a GPU: creating a context must throw, never fall back
    try {
        Context ctx(0);
        std::printf("note: a CUDA device is present, skipping the no-device check\n");
    } catch (const Error &e) {
        EXPECT(e.code == THB_ERR_CUDA);
    }
}

static Audio sine_audio(uint32_t sr, uint64_t len, double hz, float amp, uint32_t n_ch) {
    Audio a;
    a.sr = sr;
    a.len = len;
    a.n_ch = n_ch;
    a.wavs.resize(static_cast<size_t>(n_ch) * len);
    for (uint32_t ch = 0; ch < n_ch; ch++)
        for (uint64_t i = 0; i < len; i++)
            a.wavs[ch * len + i] = amp * static_cast<float>(std::sin(2.0 * M_PI * hz * (ch + 1) * i / sr));
    return a;
}

static void test_gpu_flow() {
    Context ctx(0);
    // ---- stft_works (stft.rs:173-196): impulse at n = 2, win 4 hop 2 -> |X| = 0.25 in frames 1, 2; frame 0 is zeros
    {
        TrackManager tm(ctx);
        tm.setting = SpecSetting{4.0, 2, 1, FreqScale::Linear, 0};  // sr 1000 -> win 4, hop 2, n_fft 4
        TrackList tl;
        Audio a;
        a.sr = 1000; a.len = 4; a.n_ch = 1; a.wavs = {0.0f, 0.0f, 1.0f, 0.0f};
        tl.add_tracks({7}, {a});
        tm.add_tracks(tl, {7});
        const auto sp = tm.get_spec({7, 0});
        EXPECT(sp && sp->n_frames == 3 && sp->n_bins == 3);
        if (sp) {
            const float want = 20.0f * std::log10(0.25f);
            for (int b = 0; b < 3; b++) {
                EXPECT(std::isinf(sp->at(0, b)) && sp->at(0, b) < 0);
                EXPECT(std::fabs(sp->at(1, b) - want) < 1e-3f && std::fabs(sp->at(2, b) - want) < 1e-3f);
            }
        }
        EXPECT(!tm.get_spectrogram({7, 0}));  // no image before apply_track_list_changes (mod.rs:62-72)
        const auto [ids, max_sr] = tm.apply_track_list_changes(tl);
        EXPECT(ids.size() == 1 && ids.count(7) && max_sr == 1000);
        EXPECT(tm.max_dB <= 0.0f && tm.min_dB == tm.max_dB - 100.0f);  // -inf frames -> min = max - dB_range
        const auto img = tm.get_spectrogram({7, 0});
        EXPECT(img && img->height == 3 && img->width == 3);
        if (img) {
            for (int r = 0; r < 3; r++) {
                EXPECT(img->at(r, 0) == 0);        // -inf -> 0
                EXPECT(img->at(r, 1) == 65535);    // the global maximum
            }
        }
        tm.remove_tracks(tl, tl.remove_tracks({7}));
        EXPECT(!tm.get_spectrogram({7, 0}) && !tm.get_spec({7, 0}));
    }
    // ---- TrackManager update rules on two tracks of different sample rates ----
    {
        TrackManager tm(ctx);
        TrackList tl;
        tl.add_tracks({0, 1}, {sine_audio(48000, 48000, 1000.0, 0.5f, 2), sine_audio(24000, 30000, 440.0, 0.05f, 1)});
        tm.add_tracks(tl, {0, 1});
        auto [ids, max_sr] = tm.apply_track_list_changes(tl);
        EXPECT(ids.size() == 2 && max_sr == 48000);
        const float mx0 = tm.max_dB, mn0 = tm.min_dB;
        EXPECT(mx0 < 0.0f && mx0 > -20.0f && mn0 >= mx0 - 100.0f);
        const uint32_t B0 = tm.setting.n_bins(48000), B1 = tm.setting.n_bins(24000);
        const auto i00 = tm.get_spectrogram({0, 0}), i01 = tm.get_spectrogram({0, 1}), i10 = tm.get_spectrogram({1, 0});
        EXPECT(i00 && i01 && i10);
        const auto [hop0, win0, nfft0] = tm.setting.calc_framing_params(48000);
        (void)nfft0;
        EXPECT(i00->height == B0 && i00->width == n_frames(48000, win0, hop0));
        // the 24 kHz track is drawn on the 48 kHz axis: rows past its own Nyquist exist and are zero (mod.rs:208-213)
        const auto r1 = hz_range_to_idx(FreqScale::Mel, {0.0f, 24000.0f}, 24000, B1);
        EXPECT(i10->height == r1.second - r1.first && i10->height > B1);
        bool zeros_above = true;
        for (uint64_t r = B1; r < i10->height; r++)
            for (uint64_t c = 0; c < i10->width; c++) zeros_above = zeros_above && i10->at(r, c) == 0;
        EXPECT(zeros_above);
        // nothing changed -> nothing to update (mod.rs:194-203)
        EXPECT(tm.apply_track_list_changes(tl).first.empty());
        // set_dB_range: only the quantise step runs again; the dB spectrogram is untouched
        const auto sp_before = tm.get_spec({0, 0});
        tm.set_dB_range(tl, 40.0f);
        EXPECT(tm.dB_range == 40.0f && tm.max_dB == mx0 && tm.min_dB == std::max(mn0, mx0 - 40.0f));
        const auto sp_after = tm.get_spec({0, 0});
        EXPECT(sp_before && sp_after && sp_before->dB == sp_after->dB);
        const auto i00b = tm.get_spectrogram({0, 0});
        EXPECT(i00b && i00b->px != i00->px);
        // adding a louder track moves max_dB and forces every image to be redrawn
        tl.add_tracks({2}, {sine_audio(48000, 20000, 3000.0, 0.99f, 1)});
        tm.add_tracks(tl, {2});
        EXPECT(tm.apply_track_list_changes(tl).first.size() == 3);
        EXPECT(tm.max_dB > mx0);
        // set_setting recomputes everything with the new framing
        SpecSetting lin{2048 / 48.0, 4, 1, FreqScale::Linear, 0};
        tm.set_setting(tl, lin);
        const auto il = tm.get_spectrogram({0, 0});
        EXPECT(il && il->height == 1025 && il->width == n_frames(48000, 2048, 512));
    }
    // ---- level statistics (dynamics/stats.rs:56-85): a full-scale square wave has 0 dB RMS and 0 dB peak ----
    {
        Audio a;
        a.sr = 48000; a.len = 100000; a.n_ch = 2;
        a.wavs.resize(200000);
        for (size_t i = 0; i < a.wavs.size(); i++) a.wavs[i] = (i & 1) ? 1.0f : -1.0f;
        AudioStats st = calc_stats(ctx, a);
        EXPECT(st.rms_dB == 0.0f && st.max_peak == 1.0f && st.max_peak_dB == 0.0f);
        for (float &v : a.wavs) v *= 0.5f;
        st = calc_stats(ctx, a);
        EXPECT(std::fabs(st.rms_dB - 20.0f * std::log10(0.5f)) < 1e-5f && st.max_peak == 0.5f);
    }
    // ---- Normalize / guard clipping (normalize.rs:84-110, dynamics/stats.rs:224-275) ----
    {
        AudioTrack t;
        t.original.sr = 48000; t.original.len = 3; t.original.n_ch = 2;
        t.original.wavs = {0.0f, 0.6f, -0.25f, -1.0f, 0.0f, 0.5f};
        t.original_stats = calc_stats(ctx, t.original);
        EXPECT(t.original_stats.max_peak == 1.0f && t.original_stats.max_peak_dB == 0.0f);
        t.apply_gain(ctx, 2.0f, GuardClippingMode::Clip);
        EXPECT(t.wav_before_clip[1] == 1.2f && t.wav_before_clip[3] == -2.0f && t.audio.wavs[1] == 1.0f && t.audio.wavs[3] == -1.0f);
        EXPECT(t.guard_clip_stats[0].reduction_cnt == 1 && t.guard_clip_stats[1].reduction_cnt == 1);
        EXPECT(std::fabs(t.guard_clip_stats[0].max_reduction_gain_dB - 20.0f * std::log10(1.0f / 1.2f)) < 1e-5f);
        EXPECT(std::fabs(t.guard_clip_stats[1].max_reduction_gain_dB - 20.0f * std::log10(0.5f)) < 1e-5f);
        EXPECT(t.stats.max_peak == 1.0f && t.global_gain == 1.0f);
        t.apply_gain(ctx, 2.0f, GuardClippingMode::ReduceGlobalLevel);
        EXPECT(t.global_gain == 0.5f && t.audio.wavs == t.original.wavs && t.wav_before_clip.empty());
        EXPECT(t.guard_clip_stats[0].reduction_cnt == 0 && std::fabs(t.guard_clip_stats[1].max_reduction_gain_dB + 6.0206f) < 1e-4f);
        // PeakdB(-6) on a track whose peak is 0 dB: gain 10^(-6/20), nothing to guard
        t.normalize(ctx, NormalizeTarget{NormalizeTarget::PeakdB, -6.0f}, GuardClippingMode::Clip);
        EXPECT(std::fabs(t.stats.max_peak_dB + 6.0f) < 1e-4f && t.guard_clip_stats[0].reduction_cnt == 0 && t.guard_clip_stats[0].max_reduction_gain_dB == 0.0f);
        t.normalize(ctx, NormalizeTarget{}, GuardClippingMode::Limiter);   // Off: unit gain restores the original, any mode
        EXPECT(t.audio.wavs == t.original.wavs && t.global_gain == 1.0f);
        bool threw = false;
        try { t.apply_gain(ctx, 2.0f, GuardClippingMode::Limiter); } catch (const Error &) { threw = true; }
        EXPECT(threw);
    }
    // ---- spectrogram tile KATs (render_tiles.rs:435-471) ----
    {
        RenderTileCache cache(ctx);
        cache.set_colormap({0, 0, 0, 255, 255, 0, 0, 255});
        EXPECT(cache.spectrogram_revision == 2);
        cache.set_colormap({1, 2, 3});                      // ignored, but the revision moves
        EXPECT(cache.colormap_rgba.size() == 8 && cache.spectrogram_revision == 3);
        const float spec2[2] = {0.0f, 0.0f};                // (T = 1, B = 2): only the shape matters, the image is put below
        check(thb_spec_put(ctx.get(), 77, 0, 48000, THB_FREQ_LINEAR, spec2, 1, 2), ctx.get());
        const uint16_t img[2] = {0, 65535};                 // (H = 2, W = 1)
        check(thb_img_put(ctx.get(), 77, 0, img, 2, 1), ctx.get());
        const std::vector<uint8_t> t = cache.spectrogram_tile(77, 0, 0, 0, 0, 0);
        EXPECT(t.size() == 48);
        uint64_t rev; uint32_t w, h;
        std::memcpy(&rev, t.data(), 8); std::memcpy(&w, t.data() + 8, 4); std::memcpy(&h, t.data() + 12, 4);
        EXPECT(rev == 3 && w == 1 && h == 2);
        const uint8_t hi[4] = {255, 0, 0, 255}, lo[4] = {0, 0, 0, 255};   // high frequencies first
        EXPECT(std::memcmp(t.data() + 40, hi, 4) == 0 && std::memcmp(t.data() + 44, lo, 4) == 0);
        EXPECT(cache.spectrogram_tile(77, 0, 0, 0, 3, 0).size() == 40);   // a tile past the image: header only
        check(thb_release(ctx.get(), 77, 0), ctx.get());
    }
    // ---- waveform tile KATs (render_tiles.rs:408-433) ----
    {
        const float wav[8] = {-1.0f, 0.5f, 0.25f, -0.75f, 0.1f, 0.2f, -0.3f, 0.9f};
        const std::vector<uint8_t> t = encode_waveform_tile(ctx, wav, 8, 1, 1, 0);
        EXPECT(t.size() == 24 + 12 * 4);
        uint64_t rev; uint32_t bins, spb, idx;
        std::memcpy(&rev, t.data(), 8); std::memcpy(&bins, t.data() + 8, 4); std::memcpy(&spb, t.data() + 12, 4); std::memcpy(&idx, t.data() + 16, 4);
        EXPECT(rev == 1 && bins == 4 && spb == 2 && idx == 0);
        float f[12];
        std::memcpy(f, t.data() + 24, 48);
        EXPECT(f[0] == -1.0f && f[1] == 0.5f && f[3] == -0.75f && f[4] == 0.25f && f[9] == -0.3f && f[10] == 0.9f);
        EXPECT(std::fabs(f[2] - (-0.25f)) < 1e-6f);
        // a tile past the end is empty (render_tiles.rs:240-244)
        EXPECT(encode_waveform_tile(ctx, wav, 8, 1, 1, 5).size() == 24 || encode_waveform_tile(ctx, wav, 8, 1, 1, 5).empty());
    }
}

// get_waveform_tile is served from concurrent IPC threads under read locks (lib.rs:343-389).  The library's tile
// readers take the context shared and run each call on its own stream: eight host threads asking for tiles of eight
// channels must finish in clearly less wall-clock time than the same calls made one after the other, and every answer
// must be the one a single thread gets.  (Host PCM: the channels are uploaded once by the first call, then cached.)
static void test_tile_readers_overlap() {
    Context ctx(0);
    constexpr int kThreads = 8, kCalls = 300;
    const uint64_t len = 1u << 22;  // level 11: two tiles of 2 M samples each (8 MB per call)
    std::vector<std::vector<float>> wavs(kThreads, std::vector<float>(len));
    for (int c = 0; c < kThreads; c++)
        for (uint64_t i = 0; i < len; i++) wavs[c][i] = static_cast<float>(static_cast<int>((i * 2654435761u + c * 97u) % 20001u) - 10000) / 16384.0f;
    std::vector<std::vector<uint8_t>> want(kThreads);
    for (int c = 0; c < kThreads; c++) want[c] = encode_waveform_tile(ctx, wavs[c].data(), len, 1, 11, 1);  // uploads + caches
    auto work = [&](int c, int *bad) {
        for (int i = 0; i < kCalls; i++)
            if (encode_waveform_tile(ctx, wavs[c].data(), len, 1, 11, 1) != want[c]) ++*bad;
    };
    int bad = 0;
    work(0, &bad);  // warm-up
    // Wall clock on a shared box: up to three attempts, the bar must be met once (a context-wide lock fails all three).
    bool overlapped = false;
    for (int attempt = 0; attempt < 3 && !overlapped; attempt++) {
        const auto t0 = std::chrono::steady_clock::now();
        for (int c = 0; c < kThreads; c++) work(c, &bad);
        const auto t1 = std::chrono::steady_clock::now();
        std::vector<std::thread> ths;
        std::vector<int> bads(kThreads, 0);
        for (int c = 0; c < kThreads; c++) ths.emplace_back(work, c, &bads[c]);
        for (auto &t : ths) t.join();
        const auto t2 = std::chrono::steady_clock::now();
        for (int b : bads) bad += b;
        const double serial = std::chrono::duration<double>(t1 - t0).count(), parallel = std::chrono::duration<double>(t2 - t1).count();
        std::printf("tile readers: %d x %d calls one after the other %.1f ms (%.1f us / tile), from %d threads %.1f ms (%.1f us / tile): %.2fx, %d wrong\n",
                    kThreads, kCalls, 1e3 * serial, 1e6 * serial / (kThreads * kCalls), kThreads, 1e3 * parallel,
                    1e6 * parallel / (kThreads * kCalls), serial / parallel, bad);
        // a context-wide lock would make the two equal (measured: 2.5x faster); under compute-sanitizer the tool itself
        // serialises the launches, so the wall-clock bar is skipped there (THB_NO_TIMING=1)
        overlapped = std::getenv("THB_NO_TIMING") || parallel < 0.8 * serial;
    }
    EXPECT(bad == 0);
    EXPECT(overlapped);
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "--cpu";
    try {
        test_framing();
        test_window_and_mel();
        if (mode == "--gpu") {
            test_gpu_flow();
            test_tile_readers_overlap();
        } else {
            test_no_device_is_loud();
        }
    } catch (const std::exception &e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf("%s: %s (%d failures)\n", mode.c_str(), g_fail ? "FAILED" : "ok", g_fail);
    return g_fail ? 1 : 0;
}
