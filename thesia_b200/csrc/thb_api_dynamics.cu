// thb_api_dynamics.cu -- level statistics and gain + guard clipping behind the C ABI.
#include "thb_ctx.hpp"

extern "C" {

int thb_channel_stats(thb_ctx *ctx, const thb_track *channels, size_t n, float *sum_squares, float *abs_max) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!channels || !sum_squares || !abs_max) return fail(ctx, THB_ERR_INVALID, "bad argument");
    Nvtx nv("thb_channel_stats");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n; i++) {
        if (!channels[i].pcm && channels[i].len) return fail(ctx, THB_ERR_INVALID, "channel %zu: pcm is NULL", i);
        if (channels[i].pcm_format > THB_PCM_I16) return fail(ctx, THB_ERR_INVALID, "channel %zu: pcm_format = %u", i, channels[i].pcm_format);
    }
    int rc = arena_begin(ctx, (sizeof(thb::TrackDesc) + 64) * n + 1024);
    if (rc) return rc;
    thb::TrackDesc *d_desc = nullptr;
    thb::TrackDesc *h = arena_push<thb::TrackDesc>(ctx, n, &d_desc);
    Scratch scratch(ctx);
    long long max_len = 0;
    for (size_t i = 0; i < n; i++) {
        const thb_track &t = channels[i];
        const size_t esz = t.pcm_format == THB_PCM_I16 ? 2 : 4;
        const void *d_pcm = t.pcm;
        if (t.len && !is_device_ptr(t.pcm)) {
            void *st = nullptr;
            CK(scratch.alloc(&st, esz * t.len + 64));
            CK(cudaMemcpyAsync(st, t.pcm, esz * t.len, cudaMemcpyHostToDevice, ctx->stream));
            d_pcm = st;
        }
        memset(&h[i], 0, sizeof(thb::TrackDesc));
        h[i].pcm = static_cast<const float *>(d_pcm);
        h[i].slice_len = static_cast<long long>(t.len);
        h[i].full_len = static_cast<long long>(t.len);
        h[i].pcm_i16 = t.pcm_format == THB_PCM_I16 ? 1 : 0;
        max_len = std::max(max_len, h[i].slice_len);
    }
    if ((rc = arena_commit(ctx))) return rc;
    const size_t chunks = static_cast<size_t>(thb::stats_chunks(max_len));
    double *d_part_ss = nullptr;
    float *d_part_mx = nullptr, *d_out = nullptr;
    CK(scratch.alloc(&d_part_ss, sizeof(double) * n * chunks));
    CK(scratch.alloc(&d_part_mx, sizeof(float) * n * chunks));
    CK(scratch.alloc(&d_out, sizeof(float) * 2 * n));
    {
        ProfScope ps(ctx, "channel_stats", 2 * static_cast<int>((n + 65534) / 65535));
        cudaError_t e = thb::launch_channel_stats(d_desc, static_cast<int>(n), max_len, d_part_ss, d_part_mx, d_out, d_out + n, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "channel_stats: %s", cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(sum_squares, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(abs_max, d_out + n, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_audio_stats(const float *sum_squares, const float *abs_max, const uint64_t *lens, size_t n_ch, thb_audio_stats_t *out) {
    if (!sum_squares || !abs_max || !lens || !out) return fail(nullptr, THB_ERR_INVALID, "bad argument");
    // stats.rs:66-79: the channels' f32 sums are added (rayon sum, f32), divided by the element count as f32
    float total = 0.0f, peak = 0.0f;
    uint64_t n_elem = 0;
    for (size_t c = 0; c < n_ch; c++) {
        total += sum_squares[c];
        peak = std::max(peak, abs_max[c]);
        n_elem += lens[c];
    }
    out->mean_squared = total / static_cast<float>(n_elem);
    out->rms_dB = 10.0f * log10f(out->mean_squared);   // dB_from_power_default (decibel.rs:95-107)
    out->max_peak = peak;
    out->max_peak_dB = 20.0f * log10f(peak);           // dB_from_amp_default
    return THB_OK;
}

float thb_normalize_gain(uint32_t target_kind, float target, double global_lufs, float rms_dB, float max_peak_dB) {
    // normalize.rs:31-44: 10f32.powf((target - stat) / 20.)
    switch (target_kind) {
    case THB_NORM_LUFS: return powf(10.0f, (target - static_cast<float>(global_lufs)) / 20.0f);
    case THB_NORM_RMS_DB: return powf(10.0f, (target - rms_dB) / 20.0f);
    case THB_NORM_PEAK_DB: return powf(10.0f, (target - max_peak_dB) / 20.0f);
    default: return 1.0f;
    }
}

int thb_apply_gain(thb_ctx *ctx, const thb_gain_channel *channels, size_t n, uint32_t mode, thb_gain_result *results) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!channels || !results) return fail(ctx, THB_ERR_INVALID, "bad argument");
    if (mode > THB_GUARD_LIMITER) return fail(ctx, THB_ERR_INVALID, "guard clipping mode = %u", mode);
    Nvtx nv("thb_apply_gain");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    // tracks = runs of equal id; order: channels that get a gain first, plain copies after
    std::map<uint64_t, int> group_of;
    std::vector<float> group_gain;
    std::vector<size_t> order;
    std::vector<int> grp(n);
    for (size_t i = 0; i < n; i++) {
        const thb_gain_channel &c = channels[i];
        if ((!c.pcm || !c.out) && c.len) return fail(ctx, THB_ERR_INVALID, "channel %zu: pcm / out is NULL", i);
        if (c.pcm_format > THB_PCM_I16) return fail(ctx, THB_ERR_INVALID, "channel %zu: pcm_format = %u", i, c.pcm_format);
        auto it = group_of.find(c.id);
        if (it == group_of.end()) {
            it = group_of.emplace(c.id, static_cast<int>(group_gain.size())).first;
            group_gain.push_back(c.gain);
        } else if (memcmp(&group_gain[it->second], &c.gain, sizeof(float)) != 0) {
            return fail(ctx, THB_ERR_INVALID, "channel %zu: the channels of track %llu carry different gains", i, (unsigned long long)c.id);
        }
        grp[i] = it->second;
    }
    auto is_copy = [&](size_t i) { return !std::isfinite(channels[i].gain) || channels[i].gain == 1.0f; };
    for (size_t i = 0; i < n; i++) if (!is_copy(i)) order.push_back(i);
    const size_t n_active = order.size();
    // a unit gain restores the original whatever the mode (track.rs:160-161); only a real gain needs the limiter
    if (n_active && mode == THB_GUARD_LIMITER)
        return fail(ctx, THB_ERR_UNSUPPORTED, "the limiter guard-clipping mode is a sequential recurrence and is not provided");
    for (size_t i = 0; i < n; i++) if (is_copy(i)) order.push_back(i);
    const size_t n_groups = group_gain.size();

    int rc = arena_begin(ctx, (sizeof(thb::GainDesc) + 64) * n + 1024);
    if (rc) return rc;
    thb::GainDesc *d_desc = nullptr;
    thb::GainDesc *h = arena_push<thb::GainDesc>(ctx, n, &d_desc);
    Scratch scratch(ctx);
    struct Back { void *host; const void *dev; size_t bytes; };
    std::vector<Back> backs;
    long long max_len_active = 0, max_len_copy = 0;
    for (size_t k = 0; k < n; k++) {
        const thb_gain_channel &c = channels[order[k]];
        const size_t esz = c.pcm_format == THB_PCM_I16 ? 2 : 4;
        const void *d_in = c.pcm;
        if (c.len && !is_device_ptr(c.pcm)) {
            void *p = nullptr;
            CK(scratch.alloc(&p, esz * c.len + 64));
            CK(cudaMemcpyAsync(p, c.pcm, esz * c.len, cudaMemcpyHostToDevice, ctx->stream));
            d_in = p;
        }
        auto dev_out = [&](float *user, float **dev) -> int {
            *dev = user;
            if (user && c.len && !is_device_ptr(user)) {
                void *p = nullptr;
                CK(scratch.alloc(&p, 4 * c.len + 64));
                backs.push_back({user, p, 4 * static_cast<size_t>(c.len)});
                *dev = static_cast<float *>(p);
            }
            return THB_OK;
        };
        float *d_out = nullptr, *d_before = nullptr;
        if ((rc = dev_out(c.out, &d_out))) return rc;
        const bool want_before = c.before_clip && mode == THB_GUARD_CLIP && k < n_active;
        if (want_before && (rc = dev_out(c.before_clip, &d_before))) return rc;
        memset(&h[k], 0, sizeof(thb::GainDesc));
        h[k].in = d_in;
        h[k].out = d_out;
        h[k].before = d_before;
        h[k].len = static_cast<long long>(c.len);
        h[k].gain = c.gain;
        h[k].group = grp[order[k]];
        h[k].pcm_i16 = c.pcm_format == THB_PCM_I16 ? 1 : 0;
        (k < n_active ? max_len_active : max_len_copy) = std::max(k < n_active ? max_len_active : max_len_copy, h[k].len);
    }
    if ((rc = arena_commit(ctx))) return rc;
    const size_t chunks = static_cast<size_t>(thb::gain_chunks(std::max(max_len_active, max_len_copy)));
    double *d_part = nullptr;
    thb::GainOut *d_outs = nullptr;
    unsigned *d_peak = nullptr;
    CK(scratch.alloc(&d_part, sizeof(double) * n * chunks));
    CK(scratch.alloc(&d_outs, sizeof(thb::GainOut) * n));
    CK(scratch.alloc(&d_peak, sizeof(unsigned) * n_groups));
    CK(cudaMemsetAsync(d_outs, 0, sizeof(thb::GainOut) * n, ctx->stream));
    CK(cudaMemsetAsync(d_peak, 0, sizeof(unsigned) * n_groups, ctx->stream));
    const int per_launch = 65535;
    if (n_active && mode == THB_GUARD_REDUCE_GLOBAL_LEVEL) {
        ProfScope ps(ctx, "gain_peak", static_cast<int>((n_active + per_launch - 1) / per_launch));
        cudaError_t e = thb::launch_gain_peak(d_desc, static_cast<int>(n_active), max_len_active, d_peak, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "gain_peak: %s", cudaGetErrorString(e));
    }
    if (n_active) {
        ProfScope ps(ctx, "gain_apply", 2 * static_cast<int>((n_active + per_launch - 1) / per_launch));
        cudaError_t e = thb::launch_gain_apply(d_desc, static_cast<int>(n_active), max_len_active, static_cast<int>(mode), d_peak,
                                               d_part, d_outs, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "gain_apply: %s", cudaGetErrorString(e));
    }
    if (n > n_active) {
        ProfScope ps(ctx, "gain_apply", 2 * static_cast<int>((n - n_active + per_launch - 1) / per_launch));
        cudaError_t e = thb::launch_gain_apply(d_desc + n_active, static_cast<int>(n - n_active), max_len_copy, 2, nullptr,
                                               d_part + n_active * chunks, d_outs + n_active, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "gain_apply (copy): %s", cudaGetErrorString(e));
    }
    for (const Back &b : backs) CK(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<thb::GainOut> h_outs(n);
    std::vector<unsigned> h_peak(n_groups);
    CK(cudaMemcpyAsync(h_outs.data(), d_outs, sizeof(thb::GainOut) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_peak.data(), d_peak, sizeof(unsigned) * n_groups, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    auto as_float = [](unsigned bits) { float f; memcpy(&f, &bits, 4); return f; };
    for (size_t k = 0; k < n; k++) {
        thb_gain_result &r = results[order[k]];
        const thb::GainOut &o = h_outs[k];
        r.global_gain = 1.0f;
        r.max_reduction_gain_dB = 0.0f;
        r.reduction_cnt = 0;
        r.sum_squares = o.sum_squares;
        r.abs_max = as_float(o.abs_max_bits);
        if (k >= n_active) {
            r.max_reduction_gain_dB = log10f(1.0f) * 20.0f;  // GlobalGain(1) of Audio::new (audio.rs:33-44)
        } else if (mode == THB_GUARD_CLIP) {
            const float peak = as_float(o.before_max_bits);   // GuardClippingStats::from_wav_before_clip (stats.rs:133-150)
            if (peak > 1.0f) {
                r.max_reduction_gain_dB = log10f(1.0f / peak) * 20.0f;
                r.reduction_cnt = o.reduction_cnt;
            }
        } else {
            const double peak = static_cast<double>(as_float(h_peak[grp[order[k]]]));  // audio.rs:146-159
            if (peak > 1.0) r.global_gain = static_cast<float>(1.0 / peak);
            r.max_reduction_gain_dB = log10f(r.global_gain) * 20.0f;  // from_global_gain (stats.rs:152-157)
        }
    }
    return THB_OK;
}

}  // extern "C"
