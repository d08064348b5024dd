// thb_api_core.cu -- context, plan cache, host parameter arithmetic, NCCL plumbing and measurement hooks of the
// extern "C" surface (include/thesia_b200.h).  The other entry points live in thb_api_spec.cu (update_specs /
// update_spec_imgs), thb_api_tiles.cu (tile readers) and thb_api_dynamics.cu (level statistics, gain).
#include <dlfcn.h>

#include "thb_ctx.hpp"

namespace thbapi {

thread_local std::string g_last_error;
NcclApi g_nccl;

bool NcclApi::load(std::string *err) {
    if (handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (handle) break;
    }
    if (!handle) {
        *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return false;
    }
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(handle, "ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(handle, "ncclCommInitRank"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(handle, "ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(handle, "ncclAllGather"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
        *err = "libnccl is missing a required symbol";
        return false;
    }
    return true;
}

int fail(thb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) {
        std::lock_guard<std::mutex> lk(ctx->err_mu);
        ctx->last_error = buf;
    }
    return code;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- measurement hooks -------------------------------------------------------------------------
ProfScope::ProfScope(thb_ctx *c, const char *name, int launches, cudaStream_t stream) : ctx(c), st(stream ? stream : c->stream) {
    ctx->launch_count += launches;
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    if (!ctx->profiling) return;
    e = &ctx->prof[name];
    e->launches += launches;
    auto get = [&]() {
        cudaEvent_t ev;
        if (!ctx->event_pool.empty()) {
            ev = ctx->event_pool.back();
            ctx->event_pool.pop_back();
        } else {
            cudaEventCreate(&ev);
        }
        return ev;
    };
    start = get();
    stop = get();
    cudaEventRecord(start, st);
}
ProfScope::~ProfScope() {
    if (!e) return;
    cudaEventRecord(stop, st);
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    e->pending.emplace_back(start, stop);
}

void prof_collect(thb_ctx *ctx) {
    for (auto &kv : ctx->prof) {
        for (auto &pr : kv.second.pending) {
            cudaEventSynchronize(pr.second);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) kv.second.total_ms += ms;
            ctx->event_pool.push_back(pr.first);
            ctx->event_pool.push_back(pr.second);
        }
        kv.second.pending.clear();
    }
}

// ---- descriptor arena --------------------------------------------------------------------------
int arena_begin(thb_ctx *ctx, size_t need) {
    ctx->arena_idx = (ctx->arena_idx + 1) % thb_ctx::kArenas;
    thb_ctx::Arena &a = ctx->arenas[ctx->arena_idx];
    if (!a.ev) CK(cudaEventCreateWithFlags(&a.ev, cudaEventDisableTiming));
    CK(cudaEventSynchronize(a.ev));  // the upload made from this mirror four calls ago has been consumed by the copy engine
    if (need > a.cap) {
        CK(cudaStreamSynchronize(ctx->stream));  // a queued kernel may still read the device copy
        if (a.h) cudaFreeHost(a.h);
        if (a.d) cudaFree(a.d);
        a.h = a.d = nullptr;
        a.cap = 0;
        size_t cap = 1 << 16;
        while (cap < need) cap <<= 1;
        CK(cudaMallocHost(reinterpret_cast<void **>(&a.h), cap));
        CK(cudaMalloc(reinterpret_cast<void **>(&a.d), cap));
        a.cap = cap;
    }
    ctx->h_arena = a.h;
    ctx->d_arena = a.d;
    ctx->arena_used = 0;
    return THB_OK;
}
int arena_commit(thb_ctx *ctx) {
    if (ctx->arena_used) {
        CK(cudaMemcpyAsync(ctx->d_arena, ctx->h_arena, ctx->arena_used, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaEventRecord(ctx->arenas[ctx->arena_idx].ev, ctx->stream));
    return THB_OK;
}

// ---- plan cache --------------------------------------------------------------------------------
template <typename T>
int upload(thb_ctx *ctx, Plan *pl, const std::vector<T> &v, const T **out) {
    void *d = nullptr;
    const size_t bytes = sizeof(T) * (v.empty() ? 1 : v.size());
    CK(cudaMalloc(&d, bytes));
    pl->allocs.push_back(d);
    if (!v.empty()) CK(cudaMemcpyAsync(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
    *out = static_cast<const T *>(d);
    return THB_OK;
}

int get_plan(thb_ctx *ctx, const thb_setting &s, uint32_t sr, const Plan **out) {
    const thb::Framing f = thb::framing_params(s, sr);
    if (f.hop == 0 || f.win < 3)
        return fail(ctx, THB_ERR_INVALID, "window of %llu samples (hop %llu) is too short", (unsigned long long)f.win,
                    (unsigned long long)f.hop);
    if (!thb::is_pow2(f.n_fft) || f.n_fft < 4 || f.n_fft > 32768)
        return fail(ctx, THB_ERR_UNSUPPORTED, "n_fft = %llu: only powers of two in [4, 32768] are supported",
                    (unsigned long long)f.n_fft);
    if (s.freq_scale > THB_FREQ_MEL) return fail(ctx, THB_ERR_INVALID, "freq_scale = %u", s.freq_scale);
    PlanKey key{sr, f.hop, f.win, f.n_fft, s.freq_scale, s.freq_scale == THB_FREQ_MEL ? s.n_mel : 0u};
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) {
        *out = it->second.get();
        return THB_OK;
    }
    auto pl = std::make_unique<Plan>();
    thb::PlanDev &d = pl->dev;
    d.hop = static_cast<int>(f.hop);
    d.win = static_cast<int>(f.win);
    d.n_fft = static_cast<int>(f.n_fft);
    d.nc = d.n_fft / 2;
    d.pad_left = (d.n_fft - d.win) / 2;
    d.n_freq = d.nc + 1;
    // DIF pass plan: radix 8 while possible, the remaining 2 or 4 last
    int L = 0;
    while ((1 << L) < d.nc) L++;
    d.n_pass = 0;
    while (L >= 3) {
        d.radix_log2[d.n_pass++] = 3;
        L -= 3;
    }
    if (L) d.radix_log2[d.n_pass++] = L;
    const std::vector<float> win = thb::normalized_hann(f.win, f.n_fft);
    int rc = upload(ctx, pl.get(), win, &d.window);
    if (rc) return rc;
    const std::vector<float> tw = thb::twiddle_table(f.n_fft);
    const float *twp = nullptr;
    rc = upload(ctx, pl.get(), tw, &twp);
    if (rc) return rc;
    d.twiddle = reinterpret_cast<const float2 *>(twp);
    d.n_mel = 0;
    d.max_band_len = 0;
    d.mel_nnz = 0;
    if (s.freq_scale == THB_FREQ_MEL) {
        const thb::MelBank mb = thb::mel_bank(sr, f.n_fft, s.n_mel);
        if (mb.n_mel == 0) return fail(ctx, THB_ERR_INVALID, "mel filterbank is empty for sr %u n_fft %d", sr, d.n_fft);
        d.n_mel = static_cast<int>(mb.n_mel);
        d.mel_nnz = static_cast<int>(mb.w.size());
        if ((rc = upload(ctx, pl.get(), mb.k0, &d.mel_k0))) return rc;
        if ((rc = upload(ctx, pl.get(), mb.ptr, &d.mel_ptr))) return rc;
        if ((rc = upload(ctx, pl.get(), mb.w, &d.mel_w))) return rc;
        for (uint32_t m = 0; m < mb.n_mel; m++)
            d.max_band_len = std::max<int>(d.max_band_len, static_cast<int>(mb.ptr[m + 1] - mb.ptr[m]));
    }
    d.n_bins = d.n_mel ? d.n_mel : d.n_freq;
    d.mi_words = d.mi_groups = d.mi_min_start = d.mi_max_reach = d.mi_direct = 0;
    d.mi_blob = nullptr;
    std::vector<uint32_t> mi_blob;
    if (d.n_mel && (d.n_fft == 2048 || d.n_fft == 1024 || d.n_fft == 512 || d.n_fft == 4096 || d.n_fft == 8192 || d.n_fft == 16384)) {
        // (the large-FFT kernel walks the same bin-major schedule out of global memory)
        thb::MelItems mi = thb::mel_items(thb::mel_bank(sr, f.n_fft, s.n_mel), 2, d.n_fft <= 2048, d.n_fft <= 1024);
        // THB_MEL_DIRECT=0|1 pins the mel schedule of the n_fft <= 2048 kernels (A/B runs); default: the cheaper one
        if (const char *e = getenv("THB_MEL_DIRECT")) mi.use_direct = mi.valid && d.n_fft <= 2048 && atoi(e) != 0;
        if (mi.valid) mi_blob = mi.blob();
        if (d.n_fft > 2048 && mi.valid) {
            // the large-FFT kernel keeps magnitudes (16 lead slots + reach) and two partial sums per slot in its FFT buffer
            const long long need = 16 + ((static_cast<long long>(mi.max_reach) + 2) & ~1ll) + 2ll * mi.n_groups * 32 + 1;
            if (mi.min_start < -15 || need > thb::stft_big_buffer_slots(d.n_fft)) mi_blob.clear();  // band-major fallback
        }
    }
    if (!mi_blob.empty()) {
        thb::MelItems mi = thb::mel_items(thb::mel_bank(sr, f.n_fft, s.n_mel), 2, d.n_fft <= 2048, d.n_fft <= 1024);
        if (const char *e = getenv("THB_MEL_DIRECT")) mi.use_direct = mi.valid && d.n_fft <= 2048 && atoi(e) != 0;
        d.mi_words = static_cast<int>(mi_blob.size());
        // the band-major schedule keeps no partial sums (no groups) and starts at the bands' own first bins
        d.mi_direct = mi.use_direct ? 1 : 0;
        d.mi_groups = mi.use_direct ? 0 : static_cast<int>(mi.n_groups);
        d.mi_min_start = mi.use_direct ? 0 : mi.min_start;
        d.mi_max_reach = static_cast<int>(mi.use_direct ? mi.direct_reach : mi.max_reach);
        if ((rc = upload(ctx, pl.get(), mi_blob, &d.mi_blob))) return rc;
    }
    d.big_wpad = nullptr;
    d.big_tw = nullptr;
    d.big_pieces = nullptr;
    d.big_w = nullptr;
    d.big_pptr = nullptr;
    d.big_n_pieces = 0;
    std::vector<float> bwpad, btw;
    std::vector<uint32_t> bpieces, bpptr;
    std::vector<float> bgw;
    if (d.n_fft == 1024 || d.n_fft == 4096 || d.n_fft == 8192 || d.n_fft == 16384) {
        // tables of the two-frame large-FFT kernel (thb_stft_big.cu): the frame is 256 R1 complex points = R1 x 16 x 16
        const int r1 = d.n_fft / 512, ncx = d.n_fft / 2;
        bwpad.assign(d.n_fft, 0.0f);
        for (int a = 0; a < d.win; a++) bwpad[a + d.pad_left] = 0.5f * win[a];
        // [R1 - 1][16] W_(16 R1)^(n2 k1), k1 = 1..R1-1; [R1][16] W_NC^(n3 k1); [16][16] W_256^(n3 k2)
        btw.resize(2 * ((r1 - 1) * 16 + r1 * 16 + 16 * 16));
        const double tau = 6.283185307179586476925286766559;
        auto put_tw = [&](size_t at, long long num, long long den) {
            const double a = -tau * static_cast<double>(num % den) / static_cast<double>(den);
            btw[2 * at] = static_cast<float>(std::cos(a));
            btw[2 * at + 1] = static_cast<float>(std::sin(a));
        };
        for (int k1 = 1; k1 < r1; k1++)
            for (int n2 = 0; n2 < 16; n2++) put_tw((k1 - 1) * 16 + n2, n2 * k1, 16 * r1);
        for (int k1 = 0; k1 < r1; k1++)
            for (int n3 = 0; n3 < 16; n3++) put_tw((r1 - 1) * 16 + k1 * 16 + n3, n3 * k1, ncx);
        for (int k2 = 0; k2 < 16; k2++)
            for (int n3 = 0; n3 < 16; n3++) put_tw((r1 - 1) * 16 + r1 * 16 + k2 * 16 + n3, n3 * k2, 256);
        if ((rc = upload(ctx, pl.get(), bwpad, &d.big_wpad))) return rc;
        const float *btwp = nullptr;
        if ((rc = upload(ctx, pl.get(), btw, &btwp))) return rc;
        d.big_tw = reinterpret_cast<const float2 *>(btwp);
        if (d.n_mel) {
            // Band-major pieces of <= 31 bins.  32 consecutive pieces form a group that one warp walks in lock step:
            // the group's weights are stored step-major ([step][lane], zero padded to the longest piece), so every
            // step is one coalesced 128-byte load.
            const thb::MelBank mb = thb::mel_bank(sr, f.n_fft, s.n_mel);
            bpptr.assign(mb.n_mel + 1, 0);
            std::vector<uint32_t> p_start, p_len, p_wofs;
            for (uint32_t m = 0; m < mb.n_mel; m++) {
                const uint32_t len = mb.ptr[m + 1] - mb.ptr[m];
                // 31, not 32: consecutive pieces of one band then start an odd number of bins apart and land on different
                // shared-memory banks when sixteen lanes read them in lock step
                for (uint32_t o = 0; o < len; o += 31) {
                    p_start.push_back(mb.k0[m] + o);
                    p_len.push_back(std::min<uint32_t>(31, len - o));
                    p_wofs.push_back(mb.ptr[m] + o);
                }
                bpptr[m + 1] = static_cast<uint32_t>(p_start.size());
            }
            const size_t n_pieces = p_start.size(), n_groups = (n_pieces + 31) / 32;
            d.big_n_pieces = static_cast<int>(n_pieces);
            bpieces.assign(n_groups * 32 + 2 * n_groups, 0);  // [piece] first bin, then per group {steps, weight offset}
            for (size_t g = 0; g < n_groups; g++) {
                uint32_t T = 0;
                for (size_t q = 32 * g; q < std::min(n_pieces, 32 * g + 32); q++) T = std::max(T, p_len[q]);
                bpieces[n_groups * 32 + 2 * g] = T;
                bpieces[n_groups * 32 + 2 * g + 1] = static_cast<uint32_t>(bgw.size());
                bgw.resize(bgw.size() + static_cast<size_t>(T) * 32, 0.0f);
                for (size_t q = 32 * g; q < std::min(n_pieces, 32 * g + 32); q++) {
                    bpieces[q] = p_start[q];
                    for (uint32_t i = 0; i < p_len[q]; i++)
                        bgw[bpieces[n_groups * 32 + 2 * g + 1] + static_cast<size_t>(i) * 32 + (q - 32 * g)] = mb.w[p_wofs[q] + i];
                }
            }
            if ((rc = upload(ctx, pl.get(), bpieces, &d.big_pieces))) return rc;
            if ((rc = upload(ctx, pl.get(), bgw, &d.big_w))) return rc;
            if ((rc = upload(ctx, pl.get(), bpptr, &d.big_pptr))) return rc;
        }
    }
    d.fast_wpad = nullptr;
    d.fast_tw = nullptr;
    std::vector<float> wpad, ftw;
    if (d.n_fft == 2048 || d.n_fft == 1024 || d.n_fft == 512) {
        // tables of the warp-register kernels (thb_stft_fast.cu / thb_stft_pair.cu: R1 = 32; thb_stft_warp.cu: R1 = 16 / 8):
        // the frame is NC = 32 R1 complex points = R1 x 32
        const int r1 = d.n_fft / 64, ncx = d.n_fft / 2;
        wpad.assign(d.n_fft, 0.0f);
        for (int a = 0; a < d.win; a++) wpad[a + d.pad_left] = 0.5f * win[a];
        ftw.resize(2 * ((r1 - 1) * 32 + 16 * 32));
        const double tau = 6.283185307179586476925286766559;
        for (int k1 = 1; k1 < r1; k1++)
            for (int lane = 0; lane < 32; lane++) {
                const double a = -tau * static_cast<double>((lane * k1) % ncx) / static_cast<double>(ncx);
                ftw[2 * ((k1 - 1) * 32 + lane)] = static_cast<float>(std::cos(a));
                ftw[2 * ((k1 - 1) * 32 + lane) + 1] = static_cast<float>(std::sin(a));
            }
        for (int j = 0; j < 16; j++)
            for (int lane = 0; lane < 32; lane++) {
                const int k1 = lane % r1;
                const int k_own = (k1 ? k1 : r1) + r1 * (31 - j);
                float c = tw[2 * (k_own % d.n_fft)], s = tw[2 * (k_own % d.n_fft) + 1];
                if (k_own == ncx) { c = -1.0f; s = 0.0f; }
                ftw[2 * ((r1 - 1) * 32 + j * 32 + lane)] = c;
                ftw[2 * ((r1 - 1) * 32 + j * 32 + lane) + 1] = s;
            }
        if ((rc = upload(ctx, pl.get(), wpad, &d.fast_wpad))) return rc;
        const float *ftwp = nullptr;
        if ((rc = upload(ctx, pl.get(), ftw, &ftwp))) return rc;
        d.fast_tw = reinterpret_cast<const float2 *>(ftwp);
    }
    CK(cudaStreamSynchronize(ctx->stream));  // the host vectors above die here
    *out = pl.get();
    ctx->plans[key] = std::move(pl);
    return THB_OK;
}

// ---- slots -------------------------------------------------------------------------------------
int slot_alloc(thb_ctx *ctx, int *slot) {
    if (ctx->free_slots.empty()) {
        const int new_cap = ctx->slot_cap ? ctx->slot_cap * 2 : 1024;
        float *nd = nullptr;
        CK(cudaMalloc(&nd, sizeof(float) * 2 * new_cap));
        cudaError_t e = thb::launch_minmax_init(nd, new_cap, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "minmax init: %s", cudaGetErrorString(e));
        if (ctx->d_slots) {
            CK(cudaMemcpyAsync(nd, ctx->d_slots, sizeof(float) * 2 * ctx->slot_cap, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            // slot pointers inside live descriptors are rebuilt per call, so moving is safe
            cudaFree(ctx->d_slots);
        }
        for (int i = new_cap - 1; i >= ctx->slot_cap; i--) ctx->free_slots.push_back(i);
        ctx->d_slots = nd;
        ctx->slot_cap = new_cap;
    }
    *slot = ctx->free_slots.back();
    ctx->free_slots.pop_back();
    return THB_OK;
}

void spec_free(thb_ctx *ctx, Spec &s) {
    if (s.d_spec) cudaFreeAsync(s.d_spec, ctx->stream);
    if (s.d_img) cudaFreeAsync(s.d_img, ctx->stream);
    s.d_spec = nullptr;
    s.d_img = nullptr;
    if (s.slot >= 0) {
        // a dead slot must not take part in the global reduce
        thb::launch_minmax_init(ctx->d_slots + 2 * s.slot, 1, ctx->stream);
        ctx->free_slots.push_back(s.slot);
        s.slot = -1;
    }
}

Spec *find_spec(thb_ctx *ctx, uint64_t id, uint32_t ch) {
    auto it = ctx->specs.find({id, ch});
    return it == ctx->specs.end() ? nullptr : &it->second;
}

int global_minmax_on_stream(thb_ctx *ctx, float dB_range) {
    cudaError_t e;
    if (ctx->nccl_comm && ctx->px.ok) {
        // one launch: local reduce, exchange over the peers' NVLink-mapped slots, clamp rules (thb_image.cu)
        ProfScope ps(ctx, "minmax_reduce", 1);
        e = thb::launch_minmax_exchange(ctx->d_slots, ctx->slot_cap, ctx->px.d_peers, ctx->px.mine, ctx->n_ranks, ctx->rank, ++ctx->px.seq,
                                        dB_range, ctx->d_range, ctx->d_send, ctx->px.d_fail, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "minmax exchange: %s", cudaGetErrorString(e));
        return THB_OK;
    }
    {
        ProfScope ps(ctx, "minmax_reduce", 2);
        e = thb::launch_minmax_reduce(ctx->d_slots, ctx->slot_cap, ctx->d_send, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "minmax reduce: %s", cudaGetErrorString(e));
        if (ctx->nccl_comm) {
            // the single collective of the path: {max, -min} under max  
            const int r = g_nccl.AllReduce(ctx->d_send, ctx->d_send, 2, kNcclFloat32, kNcclMax, ctx->nccl_comm, ctx->stream);
            if (r != 0)
                return fail(ctx, THB_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        }
        e = thb::launch_minmax_finalize(ctx->d_send, dB_range, ctx->d_range, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "minmax finalize: %s", cudaGetErrorString(e));
    }
    return THB_OK;
}

// Maps every rank's exchange slots into this process (CUDA IPC; the handles travel through one ncclAllGather) so that
// the global dB range needs one kernel and an NVLink round trip instead of NCCL's launch + protocol for 8 bytes.
// All ranks agree (ncclAllReduce(min) of a flag) on whether every mapping worked; if not, the NCCL path stays.
void peer_exchange_close(thb_ctx *ctx) {
    thb_ctx::PeerExchange &px = ctx->px;
    for (void *p : px.opened) cudaIpcCloseMemHandle(p);
    px.opened.clear();
    if (px.mine) cudaFree(px.mine);
    if (px.d_peers) cudaFree(px.d_peers);
    if (px.d_fail) cudaFree(px.d_fail);
    px = thb_ctx::PeerExchange{};
}

int peer_exchange_open(thb_ctx *ctx) {
    thb_ctx::PeerExchange &px = ctx->px;
    peer_exchange_close(ctx);
    const char *env = getenv("THB_PEER_EXCHANGE");
    const int n = ctx->n_ranks;
    if ((env && atoi(env) == 0) || n < 2 || n > thb::kExchangeMaxRanks || !g_nccl.AllGather) return THB_OK;
    int good = 1;
    unsigned char *d_handles = nullptr;
    int *d_flag = nullptr;
    std::vector<unsigned char> h_handles(static_cast<size_t>(n) * sizeof(cudaIpcMemHandle_t));
    std::vector<float4 *> peers(n, nullptr);
    CK(cudaMalloc(reinterpret_cast<void **>(&px.mine), sizeof(float4) * 2 * thb::kExchangeMaxRanks));
    CK(cudaMemset(px.mine, 0, sizeof(float4) * 2 * thb::kExchangeMaxRanks));   // sequence numbers start at 1
    CK(cudaMalloc(reinterpret_cast<void **>(&px.d_peers), sizeof(float4 *) * thb::kExchangeMaxRanks));
    CK(cudaMalloc(reinterpret_cast<void **>(&px.d_fail), sizeof(unsigned)));
    CK(cudaMemset(px.d_fail, 0, sizeof(unsigned)));
    CK(cudaMalloc(reinterpret_cast<void **>(&d_handles), h_handles.size()));
    CK(cudaMalloc(reinterpret_cast<void **>(&d_flag), sizeof(int)));
    cudaIpcMemHandle_t mine_h;
    if (cudaIpcGetMemHandle(&mine_h, px.mine) != cudaSuccess) {
        cudaGetLastError();
        good = 0;
        memset(&mine_h, 0, sizeof mine_h);
    }
    CK(cudaMemcpyAsync(d_handles + static_cast<size_t>(ctx->rank) * sizeof mine_h, &mine_h, sizeof mine_h, cudaMemcpyHostToDevice, ctx->stream));
    int r = g_nccl.AllGather(d_handles + static_cast<size_t>(ctx->rank) * sizeof mine_h, d_handles, sizeof mine_h, kNcclUint8, ctx->nccl_comm, ctx->stream);
    if (r != 0) return fail(ctx, THB_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    CK(cudaMemcpyAsync(h_handles.data(), d_handles, h_handles.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < n && good; q++) {
        if (q == ctx->rank) {
            peers[q] = px.mine;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, h_handles.data() + static_cast<size_t>(q) * sizeof h, sizeof h);
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
            break;
        }
        px.opened.push_back(p);
        peers[q] = static_cast<float4 *>(p);
    }
    // every rank must take the same path: min over ranks of `good` (also the barrier after which every mapping exists)
    CK(cudaMemcpyAsync(d_flag, &good, sizeof good, cudaMemcpyHostToDevice, ctx->stream));
    r = g_nccl.AllReduce(d_flag, d_flag, 1, kNcclInt32, kNcclMin, ctx->nccl_comm, ctx->stream);
    if (r != 0) return fail(ctx, THB_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    CK(cudaMemcpyAsync(&good, d_flag, sizeof good, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_handles);
    cudaFree(d_flag);
    if (!good) {
        peer_exchange_close(ctx);
        return THB_OK;   // NCCL path
    }
    CK(cudaMemcpy(px.d_peers, peers.data(), sizeof(float4 *) * n, cudaMemcpyHostToDevice));
    px.ok = true;
    return THB_OK;
}

uint64_t level_bytes(uint64_t len, uint32_t level) {
    const uint64_t spb = level < 41 ? (1ull << level) : (1ull << 40);
    const uint64_t bins = (len + spb - 1) / spb;
    const uint64_t tiles = (bins + 1023) / 1024;
    return tiles * 24 + bins * 12;
}

void put_u32(uint8_t *p, uint32_t v) {
    p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = (v >> 24) & 0xff;
}


}  // namespace thbapi

// =================================================================================================
extern "C" {

int thb_abi_version(void) { return THB_ABI_VERSION; }

// the calling thread's last error (tile readers run concurrently: a shared string would race); a thread that has
// not failed yet sees the context's last one
const char *thb_last_error(const thb_ctx *ctx) {
    if (!g_last_error.empty() || !ctx) return g_last_error.c_str();
    return ctx->last_error.c_str();
}

int thb_ctx_create(int device, void *cuda_stream, thb_ctx **out) {
    thb_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, THB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, THB_ERR_CUDA, "no CUDA device: thesia_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n_dev) return fail(nullptr, THB_ERR_INVALID, "device %d of %d", device, n_dev);
    CK(cudaSetDevice(device));
    ctx = new thb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return fail(nullptr, THB_ERR_CUDA, "cudaStreamCreate failed");
        }
        ctx->own_stream = true;
    }
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->h2d_ev, cudaEventDisableTiming);
    cudaMalloc(&ctx->d_send, sizeof(float) * 2);
    cudaMalloc(&ctx->d_range, sizeof(float) * 2);
    cudaMalloc(&ctx->d_range_tmp, sizeof(float) * 2);
    cudaMallocHost(&ctx->h_pinned, sizeof(float) * 64);
    // keep freed blocks cached in the stream-ordered pool: the path re-allocates the same sizes
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !ctx->d_send || !ctx->d_range || !ctx->h_pinned) {
        thb_ctx_destroy(ctx);
        return fail(nullptr, THB_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return THB_OK;
}

void thb_ctx_destroy(thb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    thb_comm_destroy(ctx);
    for (auto &kv : ctx->specs) spec_free(ctx, kv.second);
    ctx->specs.clear();
    for (void *p : ctx->env_outputs) cudaFreeAsync(p, ctx->stream);
    tiles_shutdown(ctx);
    ctx->tile_axes.clear();
    cudaStreamSynchronize(ctx->stream);
    ctx->plans.clear();  // ~Plan frees the tables
    prof_collect(ctx);
    for (cudaEvent_t ev : ctx->event_pool) cudaEventDestroy(ev);
    if (ctx->d_slots) cudaFree(ctx->d_slots);
    if (ctx->d_rescue_items) cudaFree(ctx->d_rescue_items);
    if (ctx->d_img_desc) cudaFree(ctx->d_img_desc);
    if (ctx->d_rescue_count) cudaFree(ctx->d_rescue_count);
    if (ctx->d_send) cudaFree(ctx->d_send);
    if (ctx->d_range) cudaFree(ctx->d_range);
    if (ctx->d_range_tmp) cudaFree(ctx->d_range_tmp);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (thb_ctx::Arena &a : ctx->arenas) {
        if (a.h) cudaFreeHost(a.h);
        if (a.d) cudaFree(a.d);
        if (a.ev) cudaEventDestroy(a.ev);
    }
    if (ctx->h2d_ev) cudaEventDestroy(ctx->h2d_ev);
    for (cudaEvent_t ev : ctx->stage_ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int thb_set_stream(thb_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) {
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return THB_OK;
}

int thb_synchronize(thb_ctx *ctx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_host_alloc(size_t bytes, void **out) {
    thb_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, THB_ERR_INVALID, "out is NULL");
    CK(cudaMallocHost(out, bytes ? bytes : 1));
    return THB_OK;
}
int thb_host_free(void *p) {
    thb_ctx *ctx = nullptr;
    if (p) CK(cudaFreeHost(p));
    return THB_OK;
}

// ---- host parameter arithmetic -------------------------------------------------------------------
int thb_framing_params(const thb_setting *s, uint32_t sr, uint64_t *hop, uint64_t *win, uint64_t *n_fft) {
    if (!s || s->t_overlap == 0) return fail(nullptr, THB_ERR_INVALID, "setting is NULL or t_overlap == 0");
    const thb::Framing f = thb::framing_params(*s, sr);
    if (hop) *hop = f.hop;
    if (win) *win = f.win;
    if (n_fft) *n_fft = f.n_fft;
    return THB_OK;
}

uint64_t thb_n_frames(uint64_t len, uint64_t win, uint64_t hop) { return thb::n_frames(len, win, hop); }

int thb_n_bins(const thb_setting *s, uint32_t sr, uint32_t *n_bins) {
    if (!s || !n_bins || s->t_overlap == 0) return fail(nullptr, THB_ERR_INVALID, "bad argument");
    const thb::Framing f = thb::framing_params(*s, sr);
    if (s->freq_scale == THB_FREQ_LINEAR) {
        *n_bins = static_cast<uint32_t>(f.n_fft / 2 + 1);
    } else if (s->n_mel) {
        *n_bins = s->n_mel;
    } else {
        *n_bins = thb::mel_bank(sr, f.n_fft, 0).n_mel;
    }
    return THB_OK;
}

int thb_hann_window(uint64_t win, uint64_t n_fft, float *out) {
    if (!out) return fail(nullptr, THB_ERR_INVALID, "out is NULL");
    const std::vector<float> w = thb::normalized_hann(win, n_fft);
    memcpy(out, w.data(), sizeof(float) * w.size());
    return THB_OK;
}

int thb_mel_fb(uint32_t sr, uint64_t n_fft, uint32_t n_mel, float *out, uint32_t *n_mel_out) {
    if (n_fft < 2 || (n_fft & 1)) return fail(nullptr, THB_ERR_INVALID, "n_fft must be even");
    const thb::MelBank mb = thb::mel_bank(sr, n_fft, n_mel);
    if (n_mel_out) *n_mel_out = mb.n_mel;
    if (out) {
        const std::vector<float> d = mb.dense();
        memcpy(out, d.data(), sizeof(float) * d.size());
    }
    return THB_OK;
}

int thb_mel_schedule_replay(uint32_t sr, uint64_t n_fft, uint32_t n_mel, float *out, uint32_t stats[4]) {
    if (!out || !stats || n_fft < 4) return THB_ERR_INVALID;
    const thb::MelBank mb = thb::mel_bank(sr, n_fft, n_mel);
    const thb::MelItems mi = thb::mel_items(mb, 2, n_fft <= 2048, n_fft <= 1024);
    stats[0] = mi.valid ? 1u : 0u;
    stats[1] = mi.n_groups;
    stats[2] = stats[3] = 0;
    if (!mi.valid) return THB_ERR_UNSUPPORTED;
    const size_t F = mb.n_freq, M = mb.n_mel, n_slots = static_cast<size_t>(mi.n_groups) * 32;
    std::fill(out, out + F * M, 0.0f);
    // band of every slot
    std::vector<int64_t> band_of(2 * n_slots, -1);
    for (size_t r = 0; r < mi.gk.size(); r++)
        for (uint32_t j = 0; j < mi.gk[r]; j++)
            for (uint32_t l = 0; l < 32; l++) {
                const uint32_t pid = mi.goff[(static_cast<size_t>(mi.gbase[r]) + j) * 32 + l];
                if (pid == mi.zero_slot) continue;
                if (pid >= 2 * n_slots || 32 * r + l >= M || band_of[pid] >= 0) return THB_ERR_INTERNAL;  // a slot feeds one band only
                band_of[pid] = static_cast<int64_t>(32 * r + l);
            }
    // the rows-of-four byte-offset table must spell the same lists, in the same order
    for (size_t r = 0; r < mi.gk.size(); r++)
        for (uint32_t l = 0; l < 32; l++) {
            std::vector<uint32_t> a, b4;
            for (uint32_t j = 0; j < mi.gk[r]; j++) {
                const uint32_t pid = mi.goff[(static_cast<size_t>(mi.gbase[r]) + j) * 32 + l];
                if (pid != mi.zero_slot) a.push_back(pid);
            }
            for (uint32_t j = 0; j < 4 * mi.gk4[r]; j++) {
                const uint32_t off = mi.goff4[(static_cast<size_t>(mi.gbase4[r]) + j / 4) * 128 + l * 4 + j % 4];
                if (off % 8) return THB_ERR_INTERNAL;
                if (off / 8 != mi.zero_slot) b4.push_back(off / 8);
            }
            if (a != b4) return THB_ERR_INTERNAL;
        }
    // rounds padded in pairs (n_fft <= 1024, mel_direct2): an even count, both rounds of a pair with one step count
    if (n_fft <= 1024) {
        if (mi.direct_L.size() % 2) return THB_ERR_INTERNAL;
        for (size_t r = 0; r + 1 < mi.direct_L.size(); r += 2)
            if (mi.direct_L[r] != mi.direct_L[r + 1]) return THB_ERR_INTERNAL;
    }
    // the band-major schedule (n_fft <= 2048) must spell the same matrix: lane l of round r walks bins k0 + i with weight w[i][l]
    for (size_t r = 0; r < mi.direct_L.size(); r++)
        for (uint32_t l = 0; l < 32; l++) {
            const size_t m = 32 * r + l;
            for (uint32_t i = 0; i < mi.direct_L[r]; i++) {
                const float w = mi.direct_w[mi.direct_woff[r] + static_cast<size_t>(i) * 32 + l];
                const int64_t k = static_cast<int64_t>(mi.direct_k0[m < M ? m : 0]) + i;
                if (m >= M) {
                    if (w != 0.0f) return THB_ERR_INTERNAL;
                    continue;
                }
                const float want = (k >= static_cast<int64_t>(mb.k0[m]) && k < static_cast<int64_t>(mb.k0[m] + (mb.ptr[m + 1] - mb.ptr[m])))
                                       ? mb.w[mb.ptr[m] + static_cast<size_t>(k - mb.k0[m])] : 0.0f;
                if (w != want) return THB_ERR_INTERNAL;
            }
            if (m < M && mi.direct_L[r] < mb.ptr[m + 1] - mb.ptr[m]) return THB_ERR_INTERNAL;
        }
    for (uint32_t g = 0; g < mi.n_groups; g++) {
        if (mi.T[g] % 2) return THB_ERR_INTERNAL;
        stats[2] += mi.T[g];
        for (uint32_t t = 0; t < mi.T[g]; t++) {
            for (uint32_t h = 0; h < 2; h++) {
                uint32_t cnt[16] = {}, worst = 0;
                for (uint32_t l = 16 * h; l < 16 * h + 16; l++) {
                    const int64_t k = static_cast<int64_t>(mi.start[g * 32 + l]) + t;
                    worst = std::max(worst, ++cnt[((k % 16) + 16) % 16]);
                }
                stats[3] += worst - 1;
            }
            for (uint32_t l = 0; l < 32; l++) {
                const int64_t k = static_cast<int64_t>(mi.start[g * 32 + l]) + t;
                for (uint32_t side = 0; side < 2; side++) {
                    const float w = mi.w[mi.w_index(g, t, l) + side];
                    if (w == 0.0f) continue;
                    const int64_t m = band_of[side * n_slots + g * 32 + l];
                    if (m < 0 || k < 0 || k >= static_cast<int64_t>(F)) return THB_ERR_INTERNAL;  // weight that reaches no band
                    if (out[static_cast<size_t>(k) * M + m] != 0.0f) return THB_ERR_INTERNAL;       // weight applied twice
                    out[static_cast<size_t>(k) * M + m] = w;
                }
            }
        }
    }
    return THB_OK;
}

int thb_hz_range_to_idx(uint32_t freq_scale, float hz0, float hz1, uint32_t sr, uint64_t n_bins, uint64_t *i0,
                        uint64_t *i1) {
    if (!i0 || !i1 || freq_scale > THB_FREQ_MEL) return fail(nullptr, THB_ERR_INVALID, "bad argument");
    thb::hz_range_to_idx(freq_scale, hz0, hz1, sr, n_bins, i0, i1);
    return THB_OK;
}

// ---- SpectrogramAnalyzer::prepare / retain (spectrogram.rs:116-185): the plan cache ---------------------
int thb_plans_prepare(thb_ctx *ctx, const thb_setting *setting, const uint32_t *srs, size_t n) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!setting || (!srs && n)) return fail(ctx, THB_ERR_INVALID, "bad argument");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n; i++) {
        const Plan *pl = nullptr;
        int rc = get_plan(ctx, *setting, srs[i], &pl);
        if (rc) return rc;
    }
    return THB_OK;
}

int thb_plan_kernel(thb_ctx *ctx, const thb_setting *setting, uint32_t sr, uint32_t *family, uint32_t *mel_schedule) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!setting || !family) return fail(ctx, THB_ERR_INVALID, "bad argument");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const Plan *pl = nullptr;
    int rc = get_plan(ctx, *setting, sr, &pl);
    if (rc) return rc;
    const thb::PlanDev &pd = pl->dev;
    // the order of thb_spec_batch (thb_api_spec.cu)
    if (thb::stft_pair_supported(pd) && thb::stft_fast_supported(pd)) *family = THB_KERNEL_PAIR;
    else if (thb::stft_warp_supported(pd)) *family = THB_KERNEL_WARP;
    else if (thb::stft_big_supported(pd)) *family = THB_KERNEL_BIG;
    else if (thb::stft_fast_supported(pd)) *family = THB_KERNEL_FAST;
    else *family = THB_KERNEL_GENERIC;
    if (mel_schedule) *mel_schedule = !pd.n_mel ? 0u : (pd.mi_blob && pd.mi_direct ? 2u : 1u);
    return THB_OK;
}

int thb_plans_retain(thb_ctx *ctx, const thb_setting *setting, const uint32_t *srs, size_t n, size_t *n_left) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!setting || (!srs && n)) return fail(ctx, THB_ERR_INVALID, "bad argument");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    // the SrWinNfft set of the track list under this setting (track.rs construct_all_sr_win_nfft_set)
    std::vector<PlanKey> keep;
    for (size_t i = 0; i < n; i++) {
        const thb::Framing f = thb::framing_params(*setting, srs[i]);
        keep.push_back(PlanKey{srs[i], f.hop, f.win, f.n_fft, setting->freq_scale, setting->freq_scale == THB_FREQ_MEL ? setting->n_mel : 0u});
    }
    bool synced = false;
    for (auto it = ctx->plans.begin(); it != ctx->plans.end();) {
        const PlanKey &k = it->first;
        bool wanted = false;
        for (const PlanKey &w : keep)
            wanted |= k.sr == w.sr && k.win == w.win && k.n_fft == w.n_fft && k.hop == w.hop && k.freq_scale == w.freq_scale &&
                      k.n_mel_req == w.n_mel_req;
        if (wanted) {
            ++it;
            continue;
        }
        if (!synced) {  // a kernel of an earlier batch may still be reading the tables
            CK(cudaStreamSynchronize(ctx->stream));
            synced = true;
        }
        it = ctx->plans.erase(it);  // ~Plan frees the tables
    }
    if (n_left) *n_left = ctx->plans.size();
    return THB_OK;
}

// ---- NCCL ---------------------------------------------------------------------------------------
int thb_comm_unique_id(uint8_t id[128]) {
    std::string err;
    if (!id) return fail(nullptr, THB_ERR_INVALID, "id is NULL");
    if (!g_nccl.load(&err)) return fail(nullptr, THB_ERR_NCCL, "%s", err.c_str());
    NcclId nid;
    const int r = g_nccl.GetUniqueId(&nid);
    if (r != 0) return fail(nullptr, THB_ERR_NCCL, "ncclGetUniqueId: %d", r);
    memcpy(id, nid.internal, 128);
    return THB_OK;
}

int thb_comm_init(thb_ctx *ctx, int n_ranks, int rank, const uint8_t id[128]) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, THB_ERR_INVALID, "bad argument");
    WriteLock lk(ctx->mu);
    std::string err;
    if (!g_nccl.load(&err)) return fail(ctx, THB_ERR_NCCL, "%s", err.c_str());
    CK(cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) {
        g_nccl.CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    NcclId nid;
    memcpy(nid.internal, id, 128);
    const int r = g_nccl.CommInitRank(&ctx->nccl_comm, n_ranks, nid, rank);
    if (r != 0) {
        ctx->nccl_comm = nullptr;
        return fail(ctx, THB_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    }
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return peer_exchange_open(ctx);
}

int thb_comm_peer_exchange(const thb_ctx *ctx) { return ctx && ctx->px.ok ? 1 : 0; }

int thb_comm_destroy(thb_ctx *ctx) {
    if (!ctx) return THB_OK;
    peer_exchange_close(ctx);
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy(ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->n_ranks = 1;
    ctx->rank = 0;
    return THB_OK;
}

// ---- measurement ------------------------------------------------------------------------------------
int thb_profile_enable(thb_ctx *ctx, int on) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    ctx->profiling = on != 0;
    return THB_OK;
}
int thb_profile_reset(thb_ctx *ctx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    ctx->prof.clear();
    ctx->launch_count = 0;
    return THB_OK;
}
int thb_profile_get(thb_ctx *ctx, const char *kernel, double *total_ms, uint64_t *launches) {
    if (!ctx || !kernel) return fail(ctx, THB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    auto it = ctx->prof.find(kernel);
    if (total_ms) *total_ms = it == ctx->prof.end() ? 0.0 : it->second.total_ms;
    if (launches) *launches = it == ctx->prof.end() ? 0 : it->second.launches;
    return THB_OK;
}
uint64_t thb_launch_count(const thb_ctx *ctx) { return ctx ? ctx->launch_count.load() : 0; }

int thb_synth_pcm(thb_ctx *ctx, float *dev_out, uint64_t len, uint32_t sr, uint32_t track, uint32_t channel,
                  uint32_t flags) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!dev_out || sr == 0 || len >= (1ull << 32)) return fail(ctx, THB_ERR_INVALID, "bad argument");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!is_device_ptr(dev_out)) return fail(ctx, THB_ERR_INVALID, "dev_out must be device memory");
    cudaError_t e = thb::launch_synth_pcm(dev_out, len, sr, track, channel, flags, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "synth: %s", cudaGetErrorString(e));
    return THB_OK;
}

}  // extern "C"
