// thb_api_spec.cu -- TrackManager::update_specs / update_spec_imgs behind the C ABI (include/thesia_b200.h):
// thb_spec_batch and the retained (id, ch) store, the global dB range with its one collective, and the u16 images.
#include "thb_ctx.hpp"

extern "C" {

// ---- update_specs ---------------------------------------------------------------------------------
int thb_spec_batch(thb_ctx *ctx, const thb_track *tracks, size_t n, const thb_setting *setting, thb_spec_out *outs) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!tracks || !setting || setting->t_overlap == 0 || setting->f_overlap == 0 || !(setting->win_ms > 0.0))
        return fail(ctx, THB_ERR_INVALID, "tracks/setting NULL, or win_ms/t_overlap/f_overlap not positive");
    Nvtx nv("thb_spec_batch");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));

    struct Item {
        const Plan *plan;
        uint64_t total_T, f_begin, f_count, full_len;
        const float *d_pcm;
        void *staging;
        int chunk;       // H2D pipeline stage this channel's samples arrive with (0 for device-resident PCM)
        bool i16;
    };
    std::vector<Item> items(n);
    // a launch group = one analyzer plan x one PCM format x one H2D pipeline stage
    using GroupKey = std::tuple<int, const Plan *, bool>;
    std::map<GroupKey, std::vector<size_t>> groups;
    // ---- validate everything before touching device state ----
    {
        // the retained store is keyed by (id, ch): two descriptors with one key would share a spectrogram buffer and
        // a min/max slot (frame-range parts of one file belong on different ranks, thesia_b200/sharding.py)
        std::vector<std::pair<uint64_t, uint32_t>> keys(n);
        for (size_t i = 0; i < n; i++) keys[i] = {tracks[i].id, tracks[i].ch};
        std::sort(keys.begin(), keys.end());
        for (size_t i = 1; i < n; i++)
            if (keys[i] == keys[i - 1])
                return fail(ctx, THB_ERR_INVALID, "(id %llu, ch %u) appears twice in one batch", (unsigned long long)keys[i].first, keys[i].second);
    }
    for (size_t i = 0; i < n; i++) {
        const thb_track &t = tracks[i];
        if (!t.pcm) return fail(ctx, THB_ERR_INVALID, "track %zu: pcm is NULL", i);
        if (t.pcm_format > THB_PCM_I16) return fail(ctx, THB_ERR_INVALID, "track %zu: pcm_format = %u", i, t.pcm_format);
        const uint64_t full_len = t.full_len ? t.full_len : t.len;
        if (full_len < 2) return fail(ctx, THB_ERR_INVALID, "track %zu: %llu samples (need >= 2; the reference's reflect pad is undefined below that)", i, (unsigned long long)full_len);
        if (t.pcm_offset + t.len > full_len) return fail(ctx, THB_ERR_INVALID, "track %zu: slice exceeds the file", i);
        const Plan *pl = nullptr;
        int rc = get_plan(ctx, *setting, t.sr, &pl);
        if (rc) return rc;
        Item &it = items[i];
        it.plan = pl;
        it.full_len = full_len;
        it.total_T = thb::n_frames(full_len, pl->dev.win, pl->dev.hop);
        it.f_begin = t.frame_begin;
        if (it.f_begin > it.total_T) return fail(ctx, THB_ERR_INVALID, "track %zu: frame_begin beyond the file", i);
        it.f_count = t.frame_count ? t.frame_count : it.total_T - it.f_begin;
        if (it.f_begin + it.f_count > it.total_T) return fail(ctx, THB_ERR_INVALID, "track %zu: frame range beyond the file", i);
        // the slice must hold every sample the frame range touches (after reflection)
        if (it.f_count) {
            const long long W = pl->dev.win, H = pl->dev.hop, N = static_cast<long long>(full_len);
            const long long lo = static_cast<long long>(it.f_begin) * H - W / 2;
            const long long hi = static_cast<long long>(it.f_begin + it.f_count - 1) * H - W / 2 + W - 1;
            long long need_lo = lo < 0 ? 0 : lo, need_hi = hi >= N ? N - 1 : hi;
            if (lo < 0) need_hi = std::max(need_hi, std::min(N - 1, -lo));
            if (hi >= N) need_lo = std::min(need_lo, std::max(0ll, 2 * (N - 1) - hi));
            if (-lo >= N || hi >= 2 * N - 1) { need_lo = 0; need_hi = N - 1; }  // multi-wrap reflect
            if (need_lo < static_cast<long long>(t.pcm_offset) || need_hi >= static_cast<long long>(t.pcm_offset + t.len))
                return fail(ctx, THB_ERR_INVALID, "track %zu: frames [%llu,+%llu) need samples [%lld,%lld] but the slice holds [%llu,%llu)", i,
                            (unsigned long long)it.f_begin, (unsigned long long)it.f_count, need_lo, need_hi,
                            (unsigned long long)t.pcm_offset, (unsigned long long)(t.pcm_offset + t.len));
        }
        it.d_pcm = nullptr;
        it.staging = nullptr;
        it.i16 = t.pcm_format == THB_PCM_I16;
        it.chunk = 0;
    }
    // H2D pipeline: host channels are cut, in call order, into stages of about kStageBytes; stage c is copied on
    // the copy stream while the kernels of stage c - 1 run, so only the last stage's compute is exposed.
    constexpr size_t kStageBytes = size_t(256) << 20;
    constexpr int kMaxStages = 64;
    int n_stages = 0;
    {
        size_t acc = 0;
        bool open = false;
        for (size_t i = 0; i < n; i++) {
            Item &it = items[i];
            if (is_device_ptr(tracks[i].pcm)) continue;
            if (!open || (acc >= kStageBytes && n_stages < kMaxStages)) {
                n_stages++;
                acc = 0;
                open = true;
            }
            it.chunk = n_stages;  // stages are numbered from 1; 0 = no copy to wait for
            acc += tracks[i].len * (it.i16 ? 2 : 4);
        }
    }
    for (size_t i = 0; i < n; i++) groups[GroupKey{items[i].chunk, items[i].plan, items[i].i16}].push_back(i);
    while (ctx->stage_ev.size() < static_cast<size_t>(n_stages)) {
        cudaEvent_t ev = nullptr;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->stage_ev.push_back(ev);
    }
    // (per group: a small job may be cut into up to 4 * sm_count extra frame-range descriptors, see below)
    int rc = arena_begin(ctx, (sizeof(thb::TrackDesc) + 256) * 4 * n +
                                  (sizeof(thb::TrackDesc) * 4 * static_cast<size_t>(ctx->sm_count) + 512) * groups.size() + 4096);
    if (rc) return rc;

    // ---- device buffers; host PCM goes through stream-ordered staging ----
    // Everything is allocated into locals first and committed to the store only when all of it succeeded: a failed
    // call leaves ctx->specs exactly as it was (no half-initialised entries, nothing leaked).
    const bool any_host = n_stages > 0;
    Scratch staging(ctx);  // the staged PCM goes back to the pool when the call ends, whatever the exit path
    struct NewSpec {
        float *d_spec = nullptr;  // fresh buffer (null: the retained one is large enough)
        int slot = -1;            // fresh slot (-1: the entry has one)
    };
    std::vector<NewSpec> fresh(n);
    auto rollback = [&]() {
        for (NewSpec &f : fresh) {
            if (f.d_spec) cudaFreeAsync(f.d_spec, ctx->stream);
            if (f.slot >= 0) ctx->free_slots.push_back(f.slot);
            f = NewSpec{};
        }
    };
    for (size_t i = 0; i < n; i++) {
        const thb_track &t = tracks[i];
        Item &it = items[i];
        const thb::PlanDev &pd = it.plan->dev;
        const Spec *old = find_spec(ctx, t.id, t.ch);
        const size_t need = static_cast<size_t>(it.f_count) * pd.n_bins;
        cudaError_t e = cudaSuccess;
        if (!old || old->slot < 0) {
            rc = slot_alloc(ctx, &fresh[i].slot);
            if (rc) {
                rollback();
                return rc;
            }
        }
        if (!old || !old->d_spec || need > old->spec_cap)
            e = cudaMallocAsync(reinterpret_cast<void **>(&fresh[i].d_spec), sizeof(float) * (need ? need : 1), ctx->stream);
        if (e == cudaSuccess) {
            if (it.chunk == 0) {
                it.d_pcm = static_cast<const float *>(t.pcm);
            } else {
                e = staging.alloc(&it.staging, (it.i16 ? 2 : 4) * t.len + 64);
                it.d_pcm = static_cast<const float *>(it.staging);
            }
        }
        if (e != cudaSuccess) {
            rollback();
            CK(e);
        }
    }
    for (size_t i = 0; i < n; i++) {
        const thb_track &t = tracks[i];
        const Item &it = items[i];
        const thb::PlanDev &pd = it.plan->dev;
        Spec &sp = ctx->specs[{t.id, t.ch}];
        if (fresh[i].slot >= 0) sp.slot = fresh[i].slot;
        if (fresh[i].d_spec) {
            if (sp.d_spec) cudaFreeAsync(sp.d_spec, ctx->stream);
            sp.d_spec = fresh[i].d_spec;
            sp.spec_cap = static_cast<size_t>(it.f_count) * pd.n_bins;
        }
        sp.id = t.id; sp.ch = t.ch; sp.sr = t.sr;
        sp.T = it.f_count; sp.total_T = it.total_T;
        sp.B = pd.n_bins; sp.hop = pd.hop; sp.win = pd.win; sp.n_fft = pd.n_fft;
        sp.freq_scale = setting->freq_scale;
    }
    // ---- descriptors, one array per plan group ----
    // THB_STFT_KERNEL = generic | fast | pair pins one implementation (A/B measurements); default: best
    const char *force = getenv("THB_STFT_KERNEL");
    const bool want_pair = !force || !strcmp(force, "pair"), want_fast = want_pair || !strcmp(force, "fast");
    const bool want_big = !force || !strcmp(force, "big");
    const bool want_warp = !force || !strcmp(force, "warp");  // n_fft 1024 / 512: thb_stft_warp.cu
    struct Launch {
        const Plan *plan;
        thb::TrackDesc *d_desc;  // every channel of the group, whole frame range
        int count;
        long long max_frames;
        // frame-pair kernel: the interior, 8-byte aligned, even-length part of every channel; the rest (file
        // edges with reflect padding, odd leftovers, unaligned channels) as separate descriptors for the scalar kernel
        thb::TrackDesc *d_pair = nullptr, *d_edge = nullptr;
        int n_pair = 0, n_edge = 0;
        long long max_pair_frames = 0, max_edge_frames = 0;
        size_t pair_tiles = 0;
        int stage = 0;
        bool i16 = false;
        bool pair_unaligned = false;  // some channel of the group needs one load per sample (odd hop / odd start)
    };
    std::vector<Launch> launches;
    size_t max_pair_tiles = 0;
    for (auto &g : groups) {
        thb::TrackDesc *d_desc = nullptr;
        thb::TrackDesc *h = arena_push<thb::TrackDesc>(ctx, g.second.size(), &d_desc);
        Launch L{std::get<1>(g.first), d_desc, static_cast<int>(g.second.size()), 0};
        L.stage = std::get<0>(g.first);
        L.i16 = std::get<2>(g.first);
        for (size_t j = 0; j < g.second.size(); j++) {
            const size_t i = g.second[j];
            const thb_track &t = tracks[i];
            const Item &it = items[i];
            const Spec &sp = ctx->specs[{t.id, t.ch}];
            h[j].pcm = it.d_pcm;
            h[j].pcm_offset = static_cast<long long>(t.pcm_offset);
            h[j].slice_len = static_cast<long long>(t.len);
            h[j].full_len = static_cast<long long>(it.full_len);
            h[j].frame_begin = static_cast<long long>(it.f_begin);
            h[j].n_frames = static_cast<long long>(it.f_count);
            h[j].out = sp.d_spec;
            h[j].minmax = ctx->d_slots + 2 * sp.slot;
            h[j].pcm_i16 = it.i16 ? 1 : 0;
            h[j].pad_ = 0;
            L.max_frames = std::max(L.max_frames, h[j].n_frames);
        }
        const thb::PlanDev &pd = L.plan->dev;
        const bool use_pair = want_pair && thb::stft_pair_supported(pd) && thb::stft_fast_supported(pd);
        const bool use_warp = want_warp && thb::stft_warp_supported(pd);
        if (use_pair || use_warp) {
            std::vector<thb::TrackDesc> pairs, edges;
            const long long W = pd.win, H = pd.hop, half = W / 2, padl = pd.pad_left;
            for (size_t j = 0; j < g.second.size(); j++) {
                const thb::TrackDesc &f = h[j];
                const long long fb = f.frame_begin, fe = fb + f.n_frames;  // [fb, fe)
                auto ceil_div = [](long long a, long long b) { return a <= 0 ? 0 : (a + b - 1) / b; };
                long long lo = std::max({fb, ceil_div(half, H), ceil_div(f.pcm_offset + half + padl, H)});
                long long hi = fe - 1;
                const long long c2 = f.full_len - W + half, c4 = f.pcm_offset + f.slice_len - pd.n_fft + half + padl;
                hi = (c2 < 0 || c4 < 0) ? -1 : std::min({hi, c2 / H, c4 / H});
                long long cnt = hi >= lo ? hi - lo + 1 : 0;
                // the frame-pair kernel loads sample pairs: 8-byte aligned float2, or 4-byte aligned i16 pairs
                const uintptr_t addr = reinterpret_cast<uintptr_t>(f.pcm);
                const int esz = L.i16 ? 2 : 4;
                const bool aligned = (addr & (esz - 1)) == 0 && (H & 1) == 0 &&
                                     ((static_cast<long long>(addr / esz) + lo * H - half - padl - f.pcm_offset) & 1) == 0;
                // channels that miss the sample-pair rule (the 44.1 kHz default has hop 441) still run on the frame-pair
                // kernel, through its one-load-per-sample variant
                const bool usable = aligned || (addr & (esz - 1)) == 0;
                if (usable && !aligned && cnt >= 2) L.pair_unaligned = true;
                // (the n_fft 2048 frame-pair kernel wants whole pairs; the n_fft 1024 / 512 kernel masks its own tail)
                cnt = usable ? (use_pair ? (cnt & ~1ll) : cnt) : 0;
                if (cnt < 2) {
                    if (f.n_frames) edges.push_back(f);
                    continue;
                }
                thb::TrackDesc pr = f;
                pr.frame_begin = lo;
                pr.n_frames = cnt;
                pr.out = f.out + (lo - fb) * pd.n_bins;
                pairs.push_back(pr);
                L.max_pair_frames = std::max(L.max_pair_frames, cnt);

                if (lo > fb) {
                    thb::TrackDesc e = f;
                    e.n_frames = lo - fb;
                    edges.push_back(e);
                }
                if (lo + cnt < fe) {
                    thb::TrackDesc e = f;
                    e.frame_begin = lo + cnt;
                    e.n_frames = fe - (lo + cnt);
                    e.out = f.out + (lo + cnt - fb) * pd.n_bins;
                    edges.push_back(e);
                }
            }
            for (const thb::TrackDesc &e : edges) L.max_edge_frames = std::max(L.max_edge_frames, e.n_frames);
            // A small job leaves the persistent grid idle or unbalanced (C1, one 44 s file: 43 tiles of 96 frames for 148
            // SMs, each of them four steps deep for its warps).  Its descriptors are cut into frame ranges of k quarter
            // tiles (a quarter tile = one step per warp), k chosen to minimise rounds x steps = ceil(items / SMs) * k; ties
            // keep the longer ranges.  Host-side only: the kernels see more, shorter descriptors, and a frame's result
            // does not depend on which descriptor carries it (the property frame-range sharding rests on).  Measured:
            // C1's packed kernel 41.7 -> 29.6 us, two 5 s channels 68 -> 45 us.  THB_SPLIT_SMALL=0 keeps whole channels.
            {
                static const bool split_ok = !(getenv("THB_SPLIT_SMALL") && atoi(getenv("THB_SPLIT_SMALL")) == 0);
                const long long tf = use_pair ? thb::stft_pair_tile_frames() : thb::stft_warp_tile_frames(pd), unit = tf / 4;
                const size_t max_cut = 4 * static_cast<size_t>(ctx->sm_count) + pairs.size();   // what the arena was sized for
                long long best_k = 4, best_cost = 0;
                if (split_ok && !pairs.empty() && unit >= 2 && unit % 2 == 0) {
                    for (long long k = 4; k >= 1; k--) {
                        long long items = 0;
                        for (const thb::TrackDesc &pr : pairs) items += (pr.n_frames + k * unit - 1) / (k * unit);
                        if (static_cast<size_t>(items) > max_cut) break;
                        const long long cost = (items + ctx->sm_count - 1) / ctx->sm_count * k;
                        if (k == 4 || cost < best_cost) {
                            best_k = k;
                            best_cost = cost;
                        }
                    }
                }
                if (best_k < 4) {
                    const long long chunk = best_k * unit;
                    std::vector<thb::TrackDesc> cut;
                    for (const thb::TrackDesc &pr : pairs)
                        for (long long off = 0; off < pr.n_frames; off += chunk) {
                            thb::TrackDesc q = pr;
                            q.frame_begin = pr.frame_begin + off;
                            q.n_frames = std::min(chunk, pr.n_frames - off);
                            q.out = pr.out + off * pd.n_bins;
                            cut.push_back(q);
                        }
                    pairs.swap(cut);
                    L.max_pair_frames = 0;
                    for (const thb::TrackDesc &pr : pairs) L.max_pair_frames = std::max(L.max_pair_frames, pr.n_frames);
                }
            }
            L.n_pair = static_cast<int>(pairs.size());
            L.n_edge = static_cast<int>(edges.size());
            if (L.n_pair) {
                thb::TrackDesc *hp = arena_push<thb::TrackDesc>(ctx, pairs.size(), &L.d_pair);
                memcpy(hp, pairs.data(), sizeof(thb::TrackDesc) * pairs.size());
            }
            if (L.n_edge) {
                thb::TrackDesc *he = arena_push<thb::TrackDesc>(ctx, edges.size(), &L.d_edge);
                memcpy(he, edges.data(), sizeof(thb::TrackDesc) * edges.size());
            }
            const long long tf = use_pair ? thb::stft_pair_tile_frames() : thb::stft_warp_tile_frames(pd);
            L.pair_tiles = static_cast<size_t>(L.n_pair) * static_cast<size_t>((L.max_pair_frames + tf - 1) / tf);
            max_pair_tiles = std::max(max_pair_tiles, L.pair_tiles);
        }
        launches.push_back(L);
    }
    rc = arena_commit(ctx);
    if (rc) return rc;
    // The PCM copies are queued only now, AFTER the descriptor upload: that small host-to-device copy on ctx->stream
    // shares the DMA queue with them, and queued behind 14 GB of PCM it held every kernel back until the last stage had
    // landed (measured: the whole 20 ms of kernel time ran after the copies instead of under them).
    if (any_host) {
        // the staging buffers exist once ctx->stream reaches this point; the copies then run stage by stage on the
        // copy stream, each stage followed by its event
        CK(cudaEventRecord(ctx->h2d_ev, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->h2d_ev, 0));
        for (int c = 1; c <= n_stages; c++) {
            for (size_t i = 0; i < n; i++)
                if (items[i].chunk == c)
                    CK(cudaMemcpyAsync(items[i].staging, tracks[i].pcm, (items[i].i16 ? 2 : 4) * tracks[i].len,
                                       cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->stage_ev[c - 1], ctx->copy_stream));
        }
    }

    if (max_pair_tiles > ctx->rescue_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_rescue_items) cudaFree(ctx->d_rescue_items);
        if (ctx->d_rescue_count) cudaFree(ctx->d_rescue_count);
        ctx->d_rescue_items = nullptr;
        ctx->d_rescue_count = nullptr;
        size_t cap = 4096;
        while (cap < max_pair_tiles) cap <<= 1;
        CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_rescue_items), sizeof(uint2) * cap));
        CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_rescue_count), sizeof(unsigned) * (cap + 2)));
        ctx->rescue_cap = cap;
    }
    // reset the {max, -min} slots of the channels being recomputed
    for (const Launch &l : launches) {
        cudaError_t e = thb::launch_minmax_init_tracks(l.d_desc, l.count, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "minmax init: %s", cudaGetErrorString(e));
        ctx->launch_count += 1;
    }

    // ---- K1/K2/K3: per (stage, plan, format) group, in stage order (std::map order) ----
    int waited = 0;
    for (const Launch &l : launches) {
        const thb::PlanDev &pd = l.plan->dev;
        if (l.stage > waited) {  // the samples of this stage must have landed
            CK(cudaStreamWaitEvent(ctx->stream, ctx->stage_ev[l.stage - 1], 0));
            waited = l.stage;
        }
        const int chunks = (l.count + 65534) / 65535;
        cudaError_t e = cudaSuccess;
        if (l.n_pair || l.n_edge) {
            const char *kname = pd.n_mel ? "stft_mel_db" : "stft_lin_db";
            const char *ename = pd.n_mel ? "stft_mel_db_edges" : "stft_lin_db_edges";
            const bool is2048 = pd.n_fft == 2048;  // else n_fft 1024 / 512: thb_stft_warp.cu, same division of labour
            // A frame-range shard of one long file has a handful of file-edge frames (reflect padding) that only the scalar
            // kernel can do: one or two CTAs for ~100 us.  Behind the packed kernel that is pure latency -- at N = 8 the two
            // end ranks of a split file finish that much after the others.  So when the edge work is
            // that small it runs on a side stream next to the packed kernel, which leaves it one SM (grid - 1: 0.7 % of
            // its throughput).  Batches of many channels (C3: 256 edge descriptors) keep the serial order: their edge
            // kernel fills the machine on its own.
            static const bool side_ok = !(getenv("THB_EDGE_SIDE") && atoi(getenv("THB_EDGE_SIDE")) == 0);
            const long long edge_ctas = static_cast<long long>(l.n_edge) * ((l.max_edge_frames + 63) / 64);
            const bool side = side_ok && l.n_pair && l.n_edge && ctx->side_stream && edge_ctas <= 2 && ctx->sm_count > 8;
            const int pair_sms = side ? ctx->sm_count - 1 : ctx->sm_count;
            auto launch_edges = [&](cudaStream_t st) {
                ProfScope ps(ctx, ename, (l.n_edge + 65534) / 65535, st);
                return is2048 ? thb::launch_stft_fast(pd, l.d_edge, l.n_edge, l.max_edge_frames, ctx->sm_count, st)
                              : thb::launch_stft_warp_scalar(pd, l.d_edge, l.n_edge, l.max_edge_frames, ctx->sm_count, st);
            };
            if (side) {
                CK(cudaEventRecord(ctx->fork_ev, ctx->stream));
                CK(cudaStreamWaitEvent(ctx->side_stream, ctx->fork_ev, 0));
                e = launch_edges(ctx->side_stream);
                CK(cudaEventRecord(ctx->join_ev, ctx->side_stream));
            }
            if (e == cudaSuccess && l.n_pair) {
                const unsigned tf = static_cast<unsigned>(is2048 ? thb::stft_pair_tile_frames() : thb::stft_warp_tile_frames(pd));
                // [0] = rescue count, [1 .. pair_tiles] = one flag per tile
                const thb::RescueList rl{ctx->d_rescue_items, ctx->d_rescue_count, ctx->d_rescue_count + 1,
                                         static_cast<unsigned>(ctx->rescue_cap), tf,
                                         static_cast<unsigned>((l.max_pair_frames + tf - 1) / tf)};
                CK(cudaMemsetAsync(ctx->d_rescue_count, 0, sizeof(unsigned) * (l.pair_tiles + 1), ctx->stream));
                {
                    ProfScope ps(ctx, kname, 1);  // the packed kernel alone: this is the roofline kernel
                    e = is2048 ? thb::launch_stft_pair(pd, l.d_pair, l.n_pair, rl, l.i16, l.pair_unaligned, pair_sms, ctx->stream)
                               : thb::launch_stft_warp_packed(pd, l.d_pair, l.n_pair, rl, l.i16, l.pair_unaligned, pair_sms, ctx->stream);
                }
                if (e == cudaSuccess) {
                    ProfScope ps(ctx, ename, 1);
                    e = is2048 ? thb::launch_stft_fast_list(pd, l.d_pair, rl, ctx->sm_count, ctx->stream)
                               : thb::launch_stft_warp_list(pd, l.d_pair, rl, ctx->sm_count, ctx->stream);
                }
            }
            if (side) {
                CK(cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0));
            } else if (e == cudaSuccess && l.n_edge) {
                e = launch_edges(ctx->stream);
            }
        } else {
            ProfScope ps(ctx, pd.n_mel ? "stft_mel_db" : "stft_lin_db", chunks);
            if (want_big && thb::stft_big_supported(pd))
                e = thb::launch_stft_big(pd, l.d_desc, l.count, l.max_frames, ctx->sm_count, ctx->stream);
            else if (want_fast && thb::stft_fast_supported(pd))
                e = thb::launch_stft_fast(pd, l.d_desc, l.count, l.max_frames, ctx->sm_count, ctx->stream);
            else
                e = thb::launch_stft_generic(pd, l.d_desc, l.count, l.max_frames, ctx->stream);
        }
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "stft launch: %s", cudaGetErrorString(e));
    }
    // ---- outputs ----
    bool any_host_out = false;
    for (size_t i = 0; i < n; i++) {
        const Item &it = items[i];
        if (!outs) continue;
        const Spec &sp = ctx->specs[{tracks[i].id, tracks[i].ch}];
        outs[i].n_frames = sp.T;
        outs[i].total_frames = sp.total_T;
        outs[i].n_bins = sp.B;
        outs[i].hop = sp.hop;
        outs[i].win = sp.win;
        outs[i].n_fft = sp.n_fft;
        if (outs[i].spec_host) {
            const uint64_t need = sp.T * sp.B;
            if (outs[i].spec_host_cap < need) {
                cudaStreamSynchronize(ctx->stream);
                return fail(ctx, THB_ERR_SMALL_BUFFER, "track %zu: spec_host holds %llu floats, need %llu", i,
                            (unsigned long long)outs[i].spec_host_cap, (unsigned long long)need);
            }
            if (need) CK(cudaMemcpyAsync(outs[i].spec_host, sp.d_spec, sizeof(float) * need, cudaMemcpyDeviceToHost, ctx->stream));
            any_host_out = true;
        }
    }
    // Host buffers belong to the caller again when we return.
    if (any_host_out) CK(cudaStreamSynchronize(ctx->stream));
    else if (any_host) CK(cudaEventSynchronize(ctx->stage_ev[n_stages - 1]));
    return THB_OK;
}

int thb_spec_put(thb_ctx *ctx, uint64_t id, uint32_t ch, uint32_t sr, uint32_t freq_scale, const float *spec,
                 uint64_t n_frames, uint32_t n_bins) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if ((!spec && n_frames * n_bins) || freq_scale > THB_FREQ_MEL) return fail(ctx, THB_ERR_INVALID, "bad argument");
    Nvtx nv("thb_spec_put");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Spec &sp = ctx->specs[{id, ch}];
    if (sp.slot < 0) {
        int rc = slot_alloc(ctx, &sp.slot);
        if (rc) return rc;
    }
    sp.id = id; sp.ch = ch; sp.sr = sr; sp.T = n_frames; sp.total_T = n_frames; sp.B = n_bins;
    sp.hop = sp.win = sp.n_fft = 0;
    sp.freq_scale = freq_scale;
    const size_t need = static_cast<size_t>(n_frames) * n_bins;
    if (need > sp.spec_cap || !sp.d_spec) {
        if (sp.d_spec) CK(cudaFreeAsync(sp.d_spec, ctx->stream));
        sp.d_spec = nullptr;
        CK(cudaMallocAsync(reinterpret_cast<void **>(&sp.d_spec), sizeof(float) * (need ? need : 1), ctx->stream));
        sp.spec_cap = need;
    }
    if (need) CK(cudaMemcpyAsync(sp.d_spec, spec, sizeof(float) * need, cudaMemcpyDefault, ctx->stream));
    cudaError_t e = thb::launch_minmax_init(ctx->d_slots + 2 * sp.slot, 1, ctx->stream);
    if (e == cudaSuccess) {
        ProfScope ps(ctx, "minmax_array", 2);
        e = thb::launch_minmax_array(sp.d_spec, need, ctx->d_slots + 2 * sp.slot, ctx->sm_count, ctx->stream);
    }
    if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "minmax: %s", cudaGetErrorString(e));
    CK(cudaStreamSynchronize(ctx->stream));  // the caller's buffer is free again
    return THB_OK;
}

int thb_spec_read(thb_ctx *ctx, uint64_t id, uint32_t ch, float *out, uint64_t cap, uint64_t *n_frames, uint32_t *n_bins) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    if (n_frames) *n_frames = sp->T;
    if (n_bins) *n_bins = sp->B;
    if (!out) return THB_OK;
    const uint64_t need = sp->T * sp->B;
    if (cap < need) return fail(ctx, THB_ERR_SMALL_BUFFER, "need %llu floats", (unsigned long long)need);
    if (need) CK(cudaMemcpyAsync(out, sp->d_spec, sizeof(float) * need, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_spec_device_ptr(thb_ctx *ctx, uint64_t id, uint32_t ch, const float **dptr, uint64_t *n_frames, uint32_t *n_bins) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    if (dptr) *dptr = sp->d_spec;
    if (n_frames) *n_frames = sp->T;
    if (n_bins) *n_bins = sp->B;
    return THB_OK;
}

int thb_spec_minmax(thb_ctx *ctx, uint64_t id, uint32_t ch, float *mn, float *mx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_slots + 2 * sp->slot, sizeof(float) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (mx) *mx = ctx->h_pinned[0];
    if (mn) *mn = -ctx->h_pinned[1];
    return THB_OK;
}

int thb_release(thb_ctx *ctx, uint64_t id, uint32_t ch) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    auto it = ctx->specs.find({id, ch});
    if (it == ctx->specs.end()) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    spec_free(ctx, it->second);
    ctx->specs.erase(it);
    return THB_OK;
}

int thb_release_all(thb_ctx *ctx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (auto &kv : ctx->specs) spec_free(ctx, kv.second);
    ctx->specs.clear();
    return THB_OK;
}

// ---- update_spec_imgs -----------------------------------------------------------------------------

// a peer that never showed up in the NVLink exchange leaves a NaN range and this flag (thb_image.cu): surface it at the
// next point where the host waits anyway
static int exchange_check(thb_ctx *ctx) {
    if (!ctx->px.ok) return THB_OK;
    unsigned f = 0;
    CK(cudaMemcpy(&f, ctx->px.d_fail, sizeof f, cudaMemcpyDeviceToHost));
    if (f) return fail(ctx, THB_ERR_NCCL, "a rank did not reach the global-range exchange within 5 s");
    return THB_OK;
}

int thb_minmax_global(thb_ctx *ctx, float dB_range, float *min_dB, float *max_dB) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    Nvtx nv("thb_minmax_global");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_slots) {
        int s;
        int rc = slot_alloc(ctx, &s);
        if (rc) return rc;
        ctx->free_slots.push_back(s);
    }
    int rc = global_minmax_on_stream(ctx, dB_range);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_range, sizeof(float) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (min_dB) *min_dB = ctx->h_pinned[0];
    if (max_dB) *max_dB = ctx->h_pinned[1];
    return exchange_check(ctx);
}

// which spec_to_img kernel a batch of descriptors may use (thb_kernels.cuh); THB_IMG_TILE=0|1|2|3 caps it (A/B runs)
static int img_tile_mode(const thb::ImgDesc *h, size_t n) {
    int mode = 3;
    for (size_t i = 0; i < n; i++) {
        if ((h[i].pitch & 1) || (reinterpret_cast<uintptr_t>(h[i].img) & 3)) return 0;
        if ((h[i].B & 3) || (h[i].i0 & 3) || (reinterpret_cast<uintptr_t>(h[i].spec) & 15)) mode = std::min(mode, 1);
    }
    if (const char *e = getenv("THB_IMG_TILE")) mode = std::min(mode, atoi(e));
    return mode < 0 ? 0 : mode;
}

static int img_prepare(thb_ctx *ctx, Spec &sp, uint64_t H) {
    const uint64_t pitch = (sp.T + 63) & ~uint64_t(63);
    const size_t need = static_cast<size_t>(H) * pitch;
    if (need > sp.img_cap || !sp.d_img) {
        if (sp.d_img) CK(cudaFreeAsync(sp.d_img, ctx->stream));
        sp.d_img = nullptr;
        CK(cudaMallocAsync(reinterpret_cast<void **>(&sp.d_img), sizeof(uint16_t) * (need ? need : 1), ctx->stream));
        sp.img_cap = need;
    }
    sp.img_H = H;
    sp.img_pitch = pitch;
    return THB_OK;
}

int thb_spec_to_img(thb_ctx *ctx, uint64_t id, uint32_t ch, uint64_t i0, uint64_t i1, float min_dB, float max_dB,
                    uint32_t colormap_length, uint16_t *out, uint64_t cap) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!out || i1 < i0 || colormap_length == 0) return fail(ctx, THB_ERR_INVALID, "bad argument");
    Nvtx nv("thb_spec_to_img");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    const uint64_t H = i1 - i0;
    if (cap < H * sp->T) return fail(ctx, THB_ERR_SMALL_BUFFER, "need %llu pixels", (unsigned long long)(H * sp->T));
    if (H == 0 || sp->T == 0) return THB_OK;
    // a scratch image, not the retained one
    Scratch scratch(ctx);
    uint16_t *d_tmp = nullptr;
    const uint64_t pitch = (sp->T + 63) & ~uint64_t(63);
    CK(scratch.alloc(&d_tmp, sizeof(uint16_t) * H * pitch));
    int rc = arena_begin(ctx, sizeof(thb::ImgDesc) + 1024);
    if (rc) return rc;
    thb::ImgDesc *d_desc = nullptr;
    thb::ImgDesc *h = arena_push<thb::ImgDesc>(ctx, 1, &d_desc);
    h->spec = sp->d_spec; h->img = d_tmp; h->T = static_cast<long long>(sp->T); h->B = static_cast<int>(sp->B);
    h->i0 = static_cast<int>(i0); h->H = static_cast<int>(H); h->pitch = static_cast<long long>(pitch);
    float *d_rng = nullptr;
    float *h_rng = arena_push<float>(ctx, 2, &d_rng);
    h_rng[0] = min_dB; h_rng[1] = max_dB;
    if ((rc = arena_commit(ctx))) return rc;
    {
        ProfScope ps(ctx, "spec_to_img");
        cudaError_t e = thb::launch_spec_to_img(d_desc, 1, h->T, h->H, d_rng, colormap_length, img_tile_mode(h, 1), ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "spec_to_img: %s", cudaGetErrorString(e));
    }
    CK(cudaMemcpy2DAsync(out, sizeof(uint16_t) * sp->T, d_tmp, sizeof(uint16_t) * pitch, sizeof(uint16_t) * sp->T, H,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

// The quantise step of update_spec_imgs for the retained spectrograms (all, or those of `only_ids`), with the range
// the stream has in ctx->d_range.  Caller holds the write lock.
static int quantise_on_stream(thb_ctx *ctx, uint32_t colormap_length, uint32_t max_sr, const uint64_t *only_ids, size_t n_only) {
    if (max_sr == 0)
        for (auto &kv : ctx->specs) max_sr = std::max(max_sr, kv.second.sr);  // tracklist.max_sr() (track.rs:371-376)
    const size_t n = ctx->specs.size();
    int rc = THB_OK;
    std::vector<unsigned char> raw(sizeof(thb::ImgDesc) * (n ? n : 1), 0);   // (zeroed: the padding takes part in the compare)
    thb::ImgDesc *h = reinterpret_cast<thb::ImgDesc *>(raw.data());
    long long max_T = 0;
    int max_H = 0;
    size_t j = 0;
    for (auto &kv : ctx->specs) {
        Spec &sp = kv.second;
        if (only_ids) {
            bool wanted = false;
            for (size_t q = 0; q < n_only && !wanted; q++) wanted = only_ids[q] == sp.id;
            if (!wanted) continue;
        }
        uint64_t i0 = 0, i1 = 0;
        // i_freq_range = hz_range_to_idx((0, max_sr / 2), sr, n_bins)  (mod.rs:208-213)
        thb::hz_range_to_idx(sp.freq_scale, 0.0f, static_cast<float>(max_sr) / 2.0f, sp.sr, sp.B, &i0, &i1);
        if ((rc = img_prepare(ctx, sp, i1 - i0))) return rc;
        h[j].spec = sp.d_spec; h[j].img = sp.d_img; h[j].T = static_cast<long long>(sp.T); h[j].B = static_cast<int>(sp.B);
        h[j].i0 = static_cast<int>(i0); h[j].H = static_cast<int>(i1 - i0); h[j].pitch = static_cast<long long>(sp.img_pitch);
        max_T = std::max(max_T, h[j].T);
        max_H = std::max(max_H, h[j].H);
        j++;
    }
    if (!j) return THB_OK;
    raw.resize(sizeof(thb::ImgDesc) * j);
    // The descriptors live in a device buffer of their own and are uploaded only when they differ from the last call's:
    // in a loop over the same retained spectrograms the quantiser then follows the range exchange directly in the
    // stream (no host-to-device copy between them: one DMA round trip less per step, and the quantiser can be launched
    // under the exchange kernel, thb_kernels.cuh launch_pdl).
    if (!ctx->d_img_desc || raw != ctx->img_desc_host) {
        if (raw.size() > ctx->img_desc_cap) {
            CK(cudaStreamSynchronize(ctx->stream));  // a queued quantiser may still read the old buffer
            if (ctx->d_img_desc) cudaFree(ctx->d_img_desc);
            ctx->d_img_desc = nullptr;
            ctx->img_desc_cap = 0;
            ctx->img_desc_host.clear();
            size_t cap = 4096;
            while (cap < raw.size()) cap <<= 1;
            CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_img_desc), cap));
            ctx->img_desc_cap = cap;
        }
        // through the pinned arena mirror (ring of four: rewritten only after this copy has been consumed)
        if ((rc = arena_begin(ctx, raw.size() + 1024))) return rc;
        unsigned char *d_unused = nullptr;
        unsigned char *pin = arena_push<unsigned char>(ctx, raw.size(), &d_unused);
        memcpy(pin, raw.data(), raw.size());
        ctx->img_desc_host.clear();   // (stays empty if the copy cannot be queued)
        CK(cudaMemcpyAsync(ctx->d_img_desc, pin, raw.size(), cudaMemcpyHostToDevice, ctx->stream));
        ctx->arena_used = 0;          // nothing for arena_commit to copy: it only marks the mirror as in flight
        if ((rc = arena_commit(ctx))) return rc;
        ctx->img_desc_host = raw;
    }
    const thb::ImgDesc *d_desc = reinterpret_cast<const thb::ImgDesc *>(ctx->d_img_desc);
    {
        ProfScope ps(ctx, "spec_to_img", static_cast<int>((j + 65534) / 65535));
        cudaError_t e = thb::launch_spec_to_img(d_desc, static_cast<int>(j), max_T, max_H, ctx->d_range, colormap_length,
                                                img_tile_mode(h, j), ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "spec_to_img: %s", cudaGetErrorString(e));
    }
    return THB_OK;
}

int thb_update_spec_imgs(thb_ctx *ctx, float dB_range, uint32_t colormap_length, uint32_t max_sr,
                         const uint64_t *only_ids, size_t n_only, float *min_dB, float *max_dB) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (colormap_length == 0) return fail(ctx, THB_ERR_INVALID, "colormap_length == 0");
    Nvtx nv("thb_update_spec_imgs");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_slots) {
        int s;
        int rc = slot_alloc(ctx, &s);
        if (rc) return rc;
        ctx->free_slots.push_back(s);
    }
    int rc = global_minmax_on_stream(ctx, dB_range);  // the one collective of the path
    if (!rc) rc = quantise_on_stream(ctx, colormap_length, max_sr, only_ids, n_only);
    if (rc) return rc;
    if (!min_dB && !max_dB) return THB_OK;  // asynchronous: thb_range_get / thb_synchronize later
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_range, sizeof(float) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (min_dB) *min_dB = ctx->h_pinned[0];
    if (max_dB) *max_dB = ctx->h_pinned[1];
    return exchange_check(ctx);
}

int thb_update_spec_imgs_range(thb_ctx *ctx, float min_dB, float max_dB, uint32_t colormap_length, uint32_t max_sr,
                               const uint64_t *only_ids, size_t n_only) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (colormap_length == 0) return fail(ctx, THB_ERR_INVALID, "colormap_length == 0");
    Nvtx nv("thb_update_spec_imgs_range");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    // the caller's (already reduced) range goes to the device through the pinned scratch; no collective here
    CK(cudaStreamSynchronize(ctx->stream));  // h_pinned may still feed an earlier copy
    ctx->h_pinned[2] = min_dB;
    ctx->h_pinned[3] = max_dB;
    CK(cudaMemcpyAsync(ctx->d_range, ctx->h_pinned + 2, sizeof(float) * 2, cudaMemcpyHostToDevice, ctx->stream));
    int rc = quantise_on_stream(ctx, colormap_length, max_sr, only_ids, n_only);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_range_get(thb_ctx *ctx, float *min_dB, float *max_dB) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_range, sizeof(float) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (min_dB) *min_dB = ctx->h_pinned[0];
    if (max_dB) *max_dB = ctx->h_pinned[1];
    return exchange_check(ctx);
}

int thb_img_read(thb_ctx *ctx, uint64_t id, uint32_t ch, uint16_t *out, uint64_t cap, uint64_t *height, uint64_t *width) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp || !sp->d_img) return fail(ctx, THB_ERR_NOT_FOUND, "no image for (%llu, %u)", (unsigned long long)id, ch);
    if (height) *height = sp->img_H;
    if (width) *width = sp->T;
    if (!out) return THB_OK;
    if (cap < sp->img_H * sp->T) return fail(ctx, THB_ERR_SMALL_BUFFER, "need %llu pixels", (unsigned long long)(sp->img_H * sp->T));
    if (sp->img_H && sp->T)
        CK(cudaMemcpy2DAsync(out, sizeof(uint16_t) * sp->T, sp->d_img, sizeof(uint16_t) * sp->img_pitch,
                             sizeof(uint16_t) * sp->T, sp->img_H, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_img_read_batch(thb_ctx *ctx, size_t n, const uint64_t *ids, const uint32_t *chs, uint16_t *const *outs,
                       const uint64_t *caps) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!ids || !chs || !outs || !caps) return fail(ctx, THB_ERR_INVALID, "bad argument");
    Nvtx nv("thb_img_read_batch");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n; i++) {  // validate everything before the first copy
        const Spec *sp = find_spec(ctx, ids[i], chs[i]);
        if (!sp || !sp->d_img) return fail(ctx, THB_ERR_NOT_FOUND, "no image for (%llu, %u)", (unsigned long long)ids[i], chs[i]);
        if (!outs[i] || caps[i] < sp->img_H * sp->T)
            return fail(ctx, THB_ERR_SMALL_BUFFER, "image %zu: need %llu pixels", i, (unsigned long long)(sp->img_H * sp->T));
    }
    for (size_t i = 0; i < n; i++) {
        const Spec *sp = find_spec(ctx, ids[i], chs[i]);
        if (sp->img_H && sp->T)
            CK(cudaMemcpy2DAsync(outs[i], sizeof(uint16_t) * sp->T, sp->d_img, sizeof(uint16_t) * sp->img_pitch,
                                 sizeof(uint16_t) * sp->T, sp->img_H, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_img_put(thb_ctx *ctx, uint64_t id, uint32_t ch, const uint16_t *img, uint64_t height, uint64_t width) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!img && height * width) return fail(ctx, THB_ERR_INVALID, "img is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    if (sp->T != width) return fail(ctx, THB_ERR_INVALID, "image width %llu != %llu frames", (unsigned long long)width, (unsigned long long)sp->T);
    int rc = img_prepare(ctx, *sp, height);
    if (rc) return rc;
    if (height && width)
        CK(cudaMemcpy2DAsync(sp->d_img, sizeof(uint16_t) * sp->img_pitch, img, sizeof(uint16_t) * width, sizeof(uint16_t) * width, height,
                             cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_img_device_ptr(thb_ctx *ctx, uint64_t id, uint32_t ch, const uint16_t **dptr, uint64_t *height, uint64_t *width,
                       uint64_t *pitch) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp || !sp->d_img) return fail(ctx, THB_ERR_NOT_FOUND, "no image for (%llu, %u)", (unsigned long long)id, ch);
    if (dptr) *dptr = sp->d_img;
    if (height) *height = sp->img_H;
    if (width) *width = sp->T;
    if (pitch) *pitch = sp->img_pitch;
    return THB_OK;
}

}  // extern "C"
