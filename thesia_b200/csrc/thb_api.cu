// thb_api.cu -- the extern "C" surface of include/thesia_b200.h.
//
// Host-side orchestration only: plan cache (SpectrogramAnalyzer::prepare, spectrogram.rs:116-154),
// the (id, ch) -> spectrogram store (TrackManager.specs / spec_imgs, mod.rs:33-44), stream-ordered
// device memory, descriptor upload, kernel launches, the 2-float NCCL all-reduce and measurement
// hooks.  All arithmetic on samples happens in the kernels; there is no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

// the C ABI is the only thing this library exports (everything else is -fvisibility=hidden)
#pragma GCC visibility push(default)
#include "../../include/thesia_b200.h"
#pragma GCC visibility pop
#include "thb_host.hpp"
#include "thb_kernels.cuh"

namespace {

thread_local std::string g_last_error;

struct PlanKey {
    uint32_t sr;
    uint64_t hop, win, n_fft;
    uint32_t freq_scale, n_mel_req;
    bool operator<(const PlanKey &o) const {
        return std::tie(sr, hop, win, n_fft, freq_scale, n_mel_req) <
               std::tie(o.sr, o.hop, o.win, o.n_fft, o.freq_scale, o.n_mel_req);
    }
};

struct Plan {
    thb::PlanDev dev{};
    std::vector<void *> allocs;
};

struct Spec {
    uint64_t id = 0;
    uint32_t ch = 0, sr = 0;
    uint64_t T = 0, total_T = 0;
    uint32_t B = 0, hop = 0, win = 0, n_fft = 0, freq_scale = 0;
    float *d_spec = nullptr;
    size_t spec_cap = 0;  // floats
    uint16_t *d_img = nullptr;
    size_t img_cap = 0;   // u16 elements
    uint64_t img_H = 0, img_pitch = 0;
    int slot = -1;
};

struct ProfEntry {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double total_ms = 0.0;
    uint64_t launches = 0;
};

// NCCL through dlopen: the library must load (and every non-collective call must work) on a box
// without NCCL, and must share torch's copy when the host program is Python.
struct NcclId {
    char internal[128];
};
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string *err) {
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
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
            *err = "libnccl is missing a required symbol";
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
constexpr int kNcclFloat32 = 7;
constexpr int kNcclMax = 2;

}  // namespace

struct thb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // H2D staging, overlapped with compute
    std::mutex mu;
    std::string last_error;

    std::map<PlanKey, std::unique_ptr<Plan>> plans;
    std::map<std::pair<uint64_t, uint32_t>, Spec> specs;

    // {max, -min} slots, one per retained spectrogram
    float *d_slots = nullptr;
    int slot_cap = 0;
    std::vector<int> free_slots;
    float *d_send = nullptr;   // 2 floats: all-reduce buffer
    float *d_range = nullptr;  // 2 floats: {min_dB, max_dB}
    float *d_range_tmp = nullptr;
    float *h_pinned = nullptr;  // small pinned scratch (64 floats)

    // descriptor arena: pinned host mirror + device copy, re-used call after call
    unsigned char *h_arena = nullptr, *d_arena = nullptr;
    size_t arena_cap = 0, arena_used = 0;
    cudaEvent_t arena_ev = nullptr;
    cudaEvent_t h2d_ev = nullptr;
    std::vector<cudaEvent_t> stage_ev;  // one per H2D pipeline stage of thb_spec_batch

    // tiles the frame-pair STFT kernel hands back to the scalar kernel (thb_kernels.cuh RescueList)
    uint2 *d_rescue_items = nullptr;
    unsigned *d_rescue_count = nullptr;  // [0] = count, [1 ..] = one flag per (descriptor, tile)
    size_t rescue_cap = 0;

    std::vector<void *> env_outputs;  // device buffers of the last waveform call

    // resize axes of the spectrogram tiles (thb_host.hpp ResizeAxis) on the device, keyed by
    // (in_size, lod_size, origin, out_size): every tile of one tile row / column of a level shares one
    struct AxisDev {
        unsigned *start = nullptr, *size = nullptr;
        int *w = nullptr;
        unsigned n = 0, window = 0, precision = 0, first = 0, end = 0;
    };
    std::map<std::tuple<uint32_t, uint64_t, uint64_t, uint32_t>, AxisDev> tile_axes;

    void *nccl_comm = nullptr;
    int n_ranks = 1, rank = 0;

    bool profiling = false;
    std::map<std::string, ProfEntry> prof;
    std::vector<cudaEvent_t> event_pool;
    uint64_t launch_count = 0;
};

namespace {

int fail(thb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->last_error = buf;
    return code;
}

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? THB_ERR_NOMEM : THB_ERR_CUDA,        \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Stream-ordered scratch memory that goes back to the pool when the scope ends, whatever the exit path.
struct Scratch {
    thb_ctx *ctx;
    std::vector<void *> ptrs;
    explicit Scratch(thb_ctx *c) : ctx(c) {}
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    template <typename T>
    cudaError_t alloc(T **p, size_t bytes) {
        void *q = nullptr;
        const cudaError_t e = cudaMallocAsync(&q, bytes ? bytes : 1, ctx->stream);
        if (e == cudaSuccess) ptrs.push_back(q);
        *p = static_cast<T *>(q);
        return e;
    }
    ~Scratch() {
        for (void *p : ptrs) cudaFreeAsync(p, ctx->stream);
    }
};

// ---- measurement hooks -------------------------------------------------------------------------
struct ProfScope {
    thb_ctx *ctx;
    ProfEntry *e = nullptr;
    cudaEvent_t start = nullptr, stop = nullptr;
    ProfScope(thb_ctx *c, const char *name, int launches = 1) : ctx(c) {
        ctx->launch_count += launches;
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
        cudaEventRecord(start, ctx->stream);
    }
    ~ProfScope() {
        if (!e) return;
        cudaEventRecord(stop, ctx->stream);
        e->pending.emplace_back(start, stop);
    }
};

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
    CK(cudaEventSynchronize(ctx->arena_ev));  // previous upload has been consumed by the copy engine
    if (need > ctx->arena_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
        if (ctx->d_arena) cudaFree(ctx->d_arena);
        ctx->h_arena = ctx->d_arena = nullptr;
        size_t cap = 1 << 16;
        while (cap < need) cap <<= 1;
        CK(cudaMallocHost(&ctx->h_arena, cap));
        CK(cudaMalloc(&ctx->d_arena, cap));
        ctx->arena_cap = cap;
    }
    ctx->arena_used = 0;
    return THB_OK;
}
template <typename T>
T *arena_push(thb_ctx *ctx, size_t count, T **dev) {
    size_t off = (ctx->arena_used + 255) & ~size_t(255);
    ctx->arena_used = off + sizeof(T) * count;
    *dev = reinterpret_cast<T *>(ctx->d_arena + off);
    return reinterpret_cast<T *>(ctx->h_arena + off);
}
int arena_commit(thb_ctx *ctx) {
    if (ctx->arena_used) {
        CK(cudaMemcpyAsync(ctx->d_arena, ctx->h_arena, ctx->arena_used, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaEventRecord(ctx->arena_ev, ctx->stream));
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
    d.mi_words = d.mi_groups = d.mi_min_start = d.mi_max_reach = 0;
    d.mi_blob = nullptr;
    std::vector<uint32_t> mi_blob;
    if (d.n_mel && (d.n_fft == 2048 || d.n_fft == 1024 || d.n_fft == 4096 || d.n_fft == 8192 || d.n_fft == 16384)) {
        // (the large-FFT kernel walks the same bin-major schedule out of global memory)
        const thb::MelItems mi = thb::mel_items(thb::mel_bank(sr, f.n_fft, s.n_mel));
        if (mi.valid) mi_blob = mi.blob();
        if (d.n_fft != 2048 && mi.valid) {
            // the large-FFT kernel keeps magnitudes (16 lead slots + reach) and two partial sums per slot in its FFT buffer
            const long long need = 16 + ((static_cast<long long>(mi.max_reach) + 2) & ~1ll) + 2ll * mi.n_groups * 32 + 1;
            if (mi.min_start < -15 || need > thb::stft_big_buffer_slots(d.n_fft)) mi_blob.clear();  // band-major fallback
        }
    }
    if (!mi_blob.empty()) {
        const thb::MelItems mi = thb::mel_items(thb::mel_bank(sr, f.n_fft, s.n_mel));
        d.mi_words = static_cast<int>(mi_blob.size());
        d.mi_groups = static_cast<int>(mi.n_groups);
        d.mi_min_start = mi.min_start;
        d.mi_max_reach = static_cast<int>(mi.max_reach);
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
    if (d.n_fft == 2048) {
        // tables of the warp-per-frame kernel (thb_stft_fast.cu)
        wpad.assign(2048, 0.0f);
        for (int a = 0; a < d.win; a++) wpad[a + d.pad_left] = 0.5f * win[a];
        ftw.resize(2 * (31 * 32 + 16 * 32));
        const double tau = 6.283185307179586476925286766559;
        for (int k1 = 1; k1 < 32; k1++)
            for (int lane = 0; lane < 32; lane++) {
                const double a = -tau * static_cast<double>((lane * k1) % 1024) / 1024.0;
                ftw[2 * ((k1 - 1) * 32 + lane)] = static_cast<float>(std::cos(a));
                ftw[2 * ((k1 - 1) * 32 + lane) + 1] = static_cast<float>(std::sin(a));
            }
        for (int j = 0; j < 16; j++)
            for (int lane = 0; lane < 32; lane++) {
                const int k_own = (lane ? lane : 32) + 32 * (31 - j);
                float c = tw[2 * (k_own % 2048)], s = tw[2 * (k_own % 2048) + 1];
                if (k_own == 1024) { c = -1.0f; s = 0.0f; }
                ftw[2 * (31 * 32 + j * 32 + lane)] = c;
                ftw[2 * (31 * 32 + j * 32 + lane) + 1] = s;
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

uint64_t level_bytes(uint64_t len, uint32_t level) {
    const uint64_t spb = level < 41 ? (1ull << level) : (1ull << 40);
    const uint64_t bins = (len + spb - 1) / spb;
    const uint64_t tiles = (bins + 1023) / 1024;
    return tiles * 24 + bins * 12;
}

void put_u32(uint8_t *p, uint32_t v) {
    p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = (v >> 24) & 0xff;
}

}  // namespace

// =================================================================================================
extern "C" {

int thb_abi_version(void) { return THB_ABI_VERSION; }

const char *thb_last_error(const thb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

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
    cudaEventCreateWithFlags(&ctx->arena_ev, cudaEventDisableTiming);
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
    for (auto &kv : ctx->tile_axes) {
        cudaFree(kv.second.start);
        cudaFree(kv.second.size);
        cudaFree(kv.second.w);
    }
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->plans)
        for (void *p : kv.second->allocs) cudaFree(p);
    prof_collect(ctx);
    for (cudaEvent_t ev : ctx->event_pool) cudaEventDestroy(ev);
    if (ctx->d_slots) cudaFree(ctx->d_slots);
    if (ctx->d_rescue_items) cudaFree(ctx->d_rescue_items);
    if (ctx->d_rescue_count) cudaFree(ctx->d_rescue_count);
    if (ctx->d_send) cudaFree(ctx->d_send);
    if (ctx->d_range) cudaFree(ctx->d_range);
    if (ctx->d_range_tmp) cudaFree(ctx->d_range_tmp);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    if (ctx->arena_ev) cudaEventDestroy(ctx->arena_ev);
    if (ctx->h2d_ev) cudaEventDestroy(ctx->h2d_ev);
    for (cudaEvent_t ev : ctx->stage_ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int thb_set_stream(thb_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    const thb::MelItems mi = thb::mel_items(mb);
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
    for (uint32_t g = 0; g < mi.n_groups; g++) {
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

// ---- update_specs ---------------------------------------------------------------------------------
int thb_spec_batch(thb_ctx *ctx, const thb_track *tracks, size_t n, const thb_setting *setting, thb_spec_out *outs) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!tracks || !setting || setting->t_overlap == 0 || setting->f_overlap == 0 || !(setting->win_ms > 0.0))
        return fail(ctx, THB_ERR_INVALID, "tracks/setting NULL, or win_ms/t_overlap/f_overlap not positive");
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    int rc = arena_begin(ctx, (sizeof(thb::TrackDesc) + 256) * 4 * n + 4096);
    if (rc) return rc;

    // ---- device buffers; host PCM goes through stream-ordered staging ----
    const bool any_host = n_stages > 0;
    for (size_t i = 0; i < n; i++) {
        const thb_track &t = tracks[i];
        Item &it = items[i];
        Spec &sp = ctx->specs[{t.id, t.ch}];
        if (sp.slot < 0) {
            rc = slot_alloc(ctx, &sp.slot);
            if (rc) return rc;
        }
        const thb::PlanDev &pd = it.plan->dev;
        sp.id = t.id; sp.ch = t.ch; sp.sr = t.sr;
        sp.T = it.f_count; sp.total_T = it.total_T;
        sp.B = pd.n_bins; sp.hop = pd.hop; sp.win = pd.win; sp.n_fft = pd.n_fft;
        sp.freq_scale = setting->freq_scale;
        const size_t need = static_cast<size_t>(sp.T) * sp.B;
        if (need > sp.spec_cap || !sp.d_spec) {
            if (sp.d_spec) CK(cudaFreeAsync(sp.d_spec, ctx->stream));
            sp.d_spec = nullptr;
            CK(cudaMallocAsync(reinterpret_cast<void **>(&sp.d_spec), sizeof(float) * (need ? need : 1), ctx->stream));
            sp.spec_cap = need;
        }
        if (it.chunk == 0) {
            it.d_pcm = static_cast<const float *>(t.pcm);
        } else {
            CK(cudaMallocAsync(&it.staging, (it.i16 ? 2 : 4) * t.len + 64, ctx->stream));
            it.d_pcm = static_cast<const float *>(it.staging);
        }
    }
    // ---- descriptors, one array per plan group ----
    // THB_STFT_KERNEL = generic | fast | pair pins one implementation (A/B measurements); default: best
    const char *force = getenv("THB_STFT_KERNEL");
    const bool want_pair = !force || !strcmp(force, "pair"), want_fast = want_pair || !strcmp(force, "fast");
    const bool want_big = !force || !strcmp(force, "big");
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
        if (want_pair && thb::stft_pair_supported(pd) && thb::stft_fast_supported(pd)) {
            std::vector<thb::TrackDesc> pairs, edges;
            const long long W = pd.win, H = pd.hop, half = W / 2, padl = pd.pad_left;
            for (size_t j = 0; j < g.second.size(); j++) {
                const thb::TrackDesc &f = h[j];
                const long long fb = f.frame_begin, fe = fb + f.n_frames;  // [fb, fe)
                auto ceil_div = [](long long a, long long b) { return a <= 0 ? 0 : (a + b - 1) / b; };
                long long lo = std::max({fb, ceil_div(half, H), ceil_div(f.pcm_offset + half + padl, H)});
                long long hi = fe - 1;
                const long long c2 = f.full_len - W + half, c4 = f.pcm_offset + f.slice_len - 2048 + half + padl;
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
                cnt = usable ? (cnt & ~1ll) : 0;
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
            const long long tf = thb::stft_pair_tile_frames();
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
        CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_rescue_count), sizeof(unsigned) * (cap + 1)));
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
            if (l.n_pair) {
                const unsigned tf = static_cast<unsigned>(thb::stft_pair_tile_frames());
                const thb::RescueList rl{ctx->d_rescue_items, ctx->d_rescue_count, ctx->d_rescue_count + 1,
                                         static_cast<unsigned>(ctx->rescue_cap), tf,
                                         static_cast<unsigned>((l.max_pair_frames + tf - 1) / tf)};
                CK(cudaMemsetAsync(ctx->d_rescue_count, 0, sizeof(unsigned) * (l.pair_tiles + 1), ctx->stream));
                {
                    ProfScope ps(ctx, kname, 1);  // the frame-pair kernel alone: this is the roofline kernel
                    e = thb::launch_stft_pair(pd, l.d_pair, l.n_pair, rl, l.i16, l.pair_unaligned, ctx->sm_count, ctx->stream);
                }
                if (e == cudaSuccess) {
                    ProfScope ps(ctx, ename, 1);
                    e = thb::launch_stft_fast_list(pd, l.d_pair, rl, ctx->sm_count, ctx->stream);
                }
            }
            if (e == cudaSuccess && l.n_edge) {
                ProfScope ps(ctx, ename, (l.n_edge + 65534) / 65535);
                e = thb::launch_stft_fast(pd, l.d_edge, l.n_edge, l.max_edge_frames, ctx->sm_count, ctx->stream);
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
        if (it.staging) CK(cudaFreeAsync(it.staging, ctx->stream));
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
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    std::lock_guard<std::mutex> lk(ctx->mu);
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    if (dptr) *dptr = sp->d_spec;
    if (n_frames) *n_frames = sp->T;
    if (n_bins) *n_bins = sp->B;
    return THB_OK;
}

int thb_spec_minmax(thb_ctx *ctx, uint64_t id, uint32_t ch, float *mn, float *mx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    auto it = ctx->specs.find({id, ch});
    if (it == ctx->specs.end()) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    spec_free(ctx, it->second);
    ctx->specs.erase(it);
    return THB_OK;
}

int thb_release_all(thb_ctx *ctx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (auto &kv : ctx->specs) spec_free(ctx, kv.second);
    ctx->specs.clear();
    return THB_OK;
}

// ---- SpectrogramAnalyzer::prepare / retain (spectrogram.rs:116-185): the plan cache ---------------------
int thb_plans_prepare(thb_ctx *ctx, const thb_setting *setting, const uint32_t *srs, size_t n) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!setting || (!srs && n)) return fail(ctx, THB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n; i++) {
        const Plan *pl = nullptr;
        int rc = get_plan(ctx, *setting, srs[i], &pl);
        if (rc) return rc;
    }
    return THB_OK;
}

int thb_plans_retain(thb_ctx *ctx, const thb_setting *setting, const uint32_t *srs, size_t n, size_t *n_left) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!setting || (!srs && n)) return fail(ctx, THB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
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
        for (void *p : it->second->allocs) cudaFree(p);
        it = ctx->plans.erase(it);
    }
    if (n_left) *n_left = ctx->plans.size();
    return THB_OK;
}

// ---- update_spec_imgs -----------------------------------------------------------------------------
int thb_minmax_global(thb_ctx *ctx, float dB_range, float *min_dB, float *max_dB) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    return THB_OK;
}

// which spec_to_img kernel a batch of descriptors may use (thb_kernels.cuh); THB_IMG_TILE=0|1|2 caps it (A/B runs)
static int img_tile_mode(const thb::ImgDesc *h, size_t n) {
    int mode = 2;
    for (size_t i = 0; i < n; i++) {
        if ((h[i].pitch & 1) || (reinterpret_cast<uintptr_t>(h[i].img) & 3)) return 0;
        if ((h[i].B & 3) || (h[i].i0 & 3) || (reinterpret_cast<uintptr_t>(h[i].spec) & 15)) mode = 1;
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
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp) return fail(ctx, THB_ERR_NOT_FOUND, "no spectrogram for (%llu, %u)", (unsigned long long)id, ch);
    const uint64_t H = i1 - i0;
    if (cap < H * sp->T) return fail(ctx, THB_ERR_SMALL_BUFFER, "need %llu pixels", (unsigned long long)(H * sp->T));
    if (H == 0 || sp->T == 0) return THB_OK;
    // a scratch image, not the retained one
    uint16_t *d_tmp = nullptr;
    const uint64_t pitch = (sp->T + 63) & ~uint64_t(63);
    CK(cudaMallocAsync(reinterpret_cast<void **>(&d_tmp), sizeof(uint16_t) * H * pitch, ctx->stream));
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
    CK(cudaFreeAsync(d_tmp, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_update_spec_imgs(thb_ctx *ctx, float dB_range, uint32_t colormap_length, uint32_t max_sr,
                         const uint64_t *only_ids, size_t n_only, float *min_dB, float *max_dB) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (colormap_length == 0) return fail(ctx, THB_ERR_INVALID, "colormap_length == 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_slots) {
        int s;
        int rc = slot_alloc(ctx, &s);
        if (rc) return rc;
        ctx->free_slots.push_back(s);
    }
    if (max_sr == 0)
        for (auto &kv : ctx->specs) max_sr = std::max(max_sr, kv.second.sr);  // tracklist.max_sr() (track.rs:371-376)
    const size_t n = ctx->specs.size();
    int rc = arena_begin(ctx, (sizeof(thb::ImgDesc) + 64) * n + 1024);
    if (rc) return rc;
    thb::ImgDesc *d_desc = nullptr;
    thb::ImgDesc *h = n ? arena_push<thb::ImgDesc>(ctx, n, &d_desc) : nullptr;
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
    if ((rc = arena_commit(ctx))) return rc;
    if ((rc = global_minmax_on_stream(ctx, dB_range))) return rc;
    if (j) {
        ProfScope ps(ctx, "spec_to_img", static_cast<int>((j + 65534) / 65535));
        cudaError_t e = thb::launch_spec_to_img(d_desc, static_cast<int>(j), max_T, max_H, ctx->d_range, colormap_length,
                                                img_tile_mode(h, j), ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "spec_to_img: %s", cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_range, sizeof(float) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (min_dB) *min_dB = ctx->h_pinned[0];
    if (max_dB) *max_dB = ctx->h_pinned[1];
    return THB_OK;
}

int thb_img_read(thb_ctx *ctx, uint64_t id, uint32_t ch, uint16_t *out, uint64_t cap, uint64_t *height, uint64_t *width) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    std::lock_guard<std::mutex> lk(ctx->mu);
    Spec *sp = find_spec(ctx, id, ch);
    if (!sp || !sp->d_img) return fail(ctx, THB_ERR_NOT_FOUND, "no image for (%llu, %u)", (unsigned long long)id, ch);
    if (dptr) *dptr = sp->d_img;
    if (height) *height = sp->img_H;
    if (width) *width = sp->T;
    if (pitch) *pitch = sp->img_pitch;
    return THB_OK;
}

// ---- spectrogram tiles (render_tiles.rs:281-393) ----------------------------------------------------
int thb_spectrogram_tile_geometry(uint64_t height, uint64_t width, uint32_t level_x, uint32_t level_y, uint32_t tile_x,
                                  uint32_t tile_y, uint64_t geo[6]) {
    if (!geo) return fail(nullptr, THB_ERR_INVALID, "geo is NULL");
    const thb::TileGeometry g = thb::spectrogram_tile_geometry(height, width, level_x, level_y, tile_x, tile_y);
    geo[0] = g.lod_width; geo[1] = g.lod_height; geo[2] = g.origin_x; geo[3] = g.origin_y; geo[4] = g.width; geo[5] = g.height;
    return THB_OK;
}

namespace {
int get_tile_axis(thb_ctx *ctx, uint32_t in_size, uint64_t lod_size, uint64_t origin, uint32_t out_size, const thb_ctx::AxisDev **out) {
    const auto key = std::make_tuple(in_size, lod_size, origin, out_size);
    auto it = ctx->tile_axes.find(key);
    if (it != ctx->tile_axes.end()) {
        *out = &it->second;
        return THB_OK;
    }
    // render_tiles.rs:379-383: the crop box in source pixels, f64
    const double in0 = static_cast<double>(origin) * static_cast<double>(in_size) / static_cast<double>(lod_size);
    const double in1 = static_cast<double>(origin + out_size) * static_cast<double>(in_size) / static_cast<double>(lod_size);
    const thb::ResizeAxis a = thb::resize_axis(in_size, in0, in0 + (in1 - in0), out_size);
    thb_ctx::AxisDev d;
    d.n = a.n;
    d.window = a.window;
    d.precision = a.precision;
    d.first = UINT32_MAX;
    d.end = 0;
    for (uint32_t o = 0; o < a.n; o++) {
        d.first = std::min(d.first, a.start[o]);
        d.end = std::max(d.end, a.start[o] + a.size[o]);
    }
    CK(cudaMalloc(reinterpret_cast<void **>(&d.start), sizeof(unsigned) * a.n));
    CK(cudaMalloc(reinterpret_cast<void **>(&d.size), sizeof(unsigned) * a.n));
    CK(cudaMalloc(reinterpret_cast<void **>(&d.w), sizeof(int) * a.w_t.size()));
    // the vectors die with this scope: plain (staged) copies, ordered before later work on the stream
    CK(cudaMemcpyAsync(d.start, a.start.data(), sizeof(unsigned) * a.n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d.size, a.size.data(), sizeof(unsigned) * a.n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d.w, a.w_t.data(), sizeof(int) * a.w_t.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = &ctx->tile_axes.emplace(key, d).first->second;
    return THB_OK;
}
}  // namespace

int thb_spectrogram_tile_batch(thb_ctx *ctx, const uint8_t *colormap_rgba, size_t colormap_bytes, uint64_t revision,
                               thb_spec_tile_req *reqs, size_t n) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!reqs) return fail(ctx, THB_ERR_INVALID, "reqs is NULL");
    if (!colormap_rgba || colormap_bytes < 4 || colormap_bytes % 4)  // RenderTileCache::set_colormap (render_tiles.rs:80-85)
        return fail(ctx, THB_ERR_INVALID, "colormap must be a non-empty RGBA table");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    struct Work { size_t req; thb::TileGeometry g; const Spec *sp; };
    std::vector<Work> work;
    for (size_t i = 0; i < n; i++) {
        thb_spec_tile_req &r = reqs[i];
        const Spec *sp = find_spec(ctx, r.id, r.ch);
        if (!sp || !sp->d_img) return fail(ctx, THB_ERR_NOT_FOUND, "no image for (%llu, %u)", (unsigned long long)r.id, r.ch);
        const thb::TileGeometry g = thb::spectrogram_tile_geometry(sp->img_H, sp->T, r.level_x, r.level_y, r.tile_x, r.tile_y);
        r.written = 40 + static_cast<size_t>(g.width) * g.height * 4;
        if (!r.out || r.cap < r.written) {
            if (r.out) return fail(ctx, THB_ERR_SMALL_BUFFER, "tile %zu: need %zu bytes", i, r.written);
            continue;  // size query
        }
        uint8_t *o = r.out;   // header (render_tiles.rs:314-323), little endian
        memcpy(o, &revision, 8);
        const uint32_t hdr[8] = {static_cast<uint32_t>(g.width), static_cast<uint32_t>(g.height), r.level_x, r.level_y, r.tile_x, r.tile_y,
                                 static_cast<uint32_t>(g.origin_x), static_cast<uint32_t>(g.origin_y)};
        memcpy(o + 8, hdr, 32);
        if (g.width && g.height) work.push_back({i, g, sp});
    }
    if (work.empty()) return THB_OK;
    const size_t m = work.size();
    std::vector<const thb_ctx::AxisDev *> axs(m), ays(m);
    if (ctx->tile_axes.size() + 2 * m > 1024) {  // bound the cache: drop everything while nothing can be reading it
        CK(cudaStreamSynchronize(ctx->stream));
        for (auto &kv : ctx->tile_axes) {
            cudaFree(kv.second.start);
            cudaFree(kv.second.size);
            cudaFree(kv.second.w);
        }
        ctx->tile_axes.clear();
    }
    for (size_t k = 0; k < m; k++) {   // a cache miss synchronises: all lookups come before the arena is in use
        const Work &w = work[k];
        int rc = get_tile_axis(ctx, static_cast<uint32_t>(w.sp->T), w.g.lod_width, w.g.origin_x, static_cast<uint32_t>(w.g.width), &axs[k]);
        if (!rc) rc = get_tile_axis(ctx, static_cast<uint32_t>(w.sp->img_H), w.g.lod_height, w.g.origin_y, static_cast<uint32_t>(w.g.height), &ays[k]);
        if (rc) return rc;
    }
    int rc = arena_begin(ctx, (sizeof(thb::TileDesc) + 64) * m + 1024);
    if (rc) return rc;
    thb::TileDesc *d_desc = nullptr;
    thb::TileDesc *h = arena_push<thb::TileDesc>(ctx, m, &d_desc);
    size_t tmp_total = 0, out_total = 0;
    unsigned max_w = 0, max_h = 0, max_tmp_h = 0;
    for (size_t k = 0; k < m; k++) {
        const unsigned tmp_h = ays[k]->end - ays[k]->first;
        tmp_total += (static_cast<size_t>(tmp_h) * work[k].g.width + 7) & ~size_t(7);
        out_total += static_cast<size_t>(work[k].g.width) * work[k].g.height * 4;
        max_w = std::max<unsigned>(max_w, static_cast<unsigned>(work[k].g.width));
        max_h = std::max<unsigned>(max_h, static_cast<unsigned>(work[k].g.height));
        max_tmp_h = std::max(max_tmp_h, tmp_h);
    }
    Scratch scratch(ctx);
    uint16_t *d_tmp = nullptr;
    uint8_t *d_out = nullptr;
    uchar4 *d_cm = nullptr;
    CK(scratch.alloc(&d_tmp, sizeof(uint16_t) * tmp_total + 16));
    CK(scratch.alloc(&d_out, out_total + 16));
    CK(scratch.alloc(&d_cm, colormap_bytes));
    CK(cudaMemcpyAsync(d_cm, colormap_rgba, colormap_bytes, cudaMemcpyHostToDevice, ctx->stream));
    size_t tmp_off = 0, out_off = 0;
    for (size_t k = 0; k < m; k++) {
        const Work &w = work[k];
        thb::TileDesc &t = h[k];
        memset(&t, 0, sizeof t);
        t.img = w.sp->d_img;
        t.pitch = w.sp->img_pitch;
        t.x_start = axs[k]->start;
        t.x_size = axs[k]->size;
        t.wx = axs[k]->w;
        t.y_start = ays[k]->start;
        t.y_size = ays[k]->size;
        t.wy = ays[k]->w;
        t.width = static_cast<unsigned>(w.g.width);
        t.height = static_cast<unsigned>(w.g.height);
        t.y_first = ays[k]->first;
        t.tmp_h = ays[k]->end - ays[k]->first;
        t.px = axs[k]->precision;
        t.py = ays[k]->precision;
        t.tmp = d_tmp + tmp_off;
        t.out = d_out + out_off;
        tmp_off += (static_cast<size_t>(t.tmp_h) * t.width + 7) & ~size_t(7);
        out_off += static_cast<size_t>(t.width) * t.height * 4;
    }
    if ((rc = arena_commit(ctx))) return rc;
    {
        ProfScope ps(ctx, "spectrogram_tile", 2 * static_cast<int>((m + 65534) / 65535));
        cudaError_t e = thb::launch_spectrogram_tiles(d_desc, static_cast<int>(m), max_w, max_h, max_tmp_h, d_cm,
                                                      static_cast<unsigned>(colormap_bytes / 4), ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "spectrogram_tile: %s", cudaGetErrorString(e));
    }
    out_off = 0;
    for (size_t k = 0; k < m; k++) {
        const size_t bytes = static_cast<size_t>(work[k].g.width) * work[k].g.height * 4;
        CK(cudaMemcpyAsync(reqs[work[k].req].out + 40, d_out + out_off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        out_off += bytes;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_spectrogram_tile(thb_ctx *ctx, uint64_t id, uint32_t ch, const uint8_t *colormap_rgba, size_t colormap_bytes,
                         uint64_t revision, uint32_t level_x, uint32_t level_y, uint32_t tile_x, uint32_t tile_y, uint8_t *out,
                         size_t cap, size_t *written) {
    thb_spec_tile_req r{};
    r.id = id;
    r.ch = ch;
    r.level_x = level_x;
    r.level_y = level_y;
    r.tile_x = tile_x;
    r.tile_y = tile_y;
    r.out = out;
    r.cap = cap;
    const int rc = thb_spectrogram_tile_batch(ctx, colormap_rgba, colormap_bytes, revision, &r, 1);
    if (written) *written = r.written;
    return rc;
}

// ---- waveform tiles ---------------------------------------------------------------------------------
uint64_t thb_waveform_level_bytes(uint64_t len, uint32_t level) { return level_bytes(len, level); }

int thb_waveform_level_batch(thb_ctx *ctx, const thb_track *tracks, size_t n, uint64_t revision, uint32_t level,
                             uint8_t **host_out, const size_t *caps, size_t *written, const uint8_t **dev_out) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!tracks) return fail(ctx, THB_ERR_INVALID, "tracks is NULL");
    if (level > 40) return fail(ctx, THB_ERR_INVALID, "level %u", level);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (void *p : ctx->env_outputs) CK(cudaFreeAsync(p, ctx->stream));
    ctx->env_outputs.clear();
    int rc = arena_begin(ctx, (sizeof(thb::EnvDesc) + 64) * n + 1024);
    if (rc) return rc;
    thb::EnvDesc *d_desc = nullptr;
    thb::EnvDesc *h = arena_push<thb::EnvDesc>(ctx, n, &d_desc);
    std::vector<void *> staging(n, nullptr);
    std::vector<uint64_t> bytes(n);
    long long max_len = 0;
    bool any_host = false;
    for (size_t i = 0; i < n; i++) {
        const thb_track &t = tracks[i];
        if (!t.pcm && t.len) return fail(ctx, THB_ERR_INVALID, "track %zu: pcm is NULL", i);
        if (t.pcm_format != THB_PCM_F32) return fail(ctx, THB_ERR_UNSUPPORTED, "track %zu: waveform tiles take f32 PCM", i);
        bytes[i] = level_bytes(t.len, level);
        if (host_out && host_out[i] && caps && caps[i] < bytes[i])
            return fail(ctx, THB_ERR_SMALL_BUFFER, "track %zu: need %llu bytes", i, (unsigned long long)bytes[i]);
        const float *d_pcm = static_cast<const float *>(t.pcm);
        if (t.len && !is_device_ptr(t.pcm)) {
            CK(cudaMallocAsync(&staging[i], sizeof(float) * t.len + 64, ctx->stream));
            CK(cudaMemcpyAsync(staging[i], t.pcm, sizeof(float) * t.len, cudaMemcpyHostToDevice, ctx->stream));
            d_pcm = static_cast<const float *>(staging[i]);
            any_host = true;
        }
        void *d_out = nullptr;
        CK(cudaMallocAsync(&d_out, bytes[i] ? bytes[i] : 4, ctx->stream));
        ctx->env_outputs.push_back(d_out);
        h[i].pcm = d_pcm;
        h[i].len = static_cast<long long>(t.len);
        h[i].out = static_cast<uint8_t *>(d_out);
        max_len = std::max(max_len, h[i].len);
        if (dev_out) dev_out[i] = static_cast<const uint8_t *>(d_out);
        if (written) written[i] = bytes[i];
    }
    if ((rc = arena_commit(ctx))) return rc;
    {
        ProfScope ps(ctx, "envelope", static_cast<int>((n + 65534) / 65535));
        cudaError_t e = thb::launch_envelope(d_desc, static_cast<int>(n), max_len, level, revision, 0, 0, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "envelope: %s", cudaGetErrorString(e));
    }
    bool any_out = false;
    for (size_t i = 0; i < n; i++) {
        if (staging[i]) CK(cudaFreeAsync(staging[i], ctx->stream));
        if (host_out && host_out[i] && bytes[i]) {
            CK(cudaMemcpyAsync(host_out[i], ctx->env_outputs[i], bytes[i], cudaMemcpyDeviceToHost, ctx->stream));
            any_out = true;
        }
    }
    if (any_out || any_host) CK(cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_channel_stats(thb_ctx *ctx, const thb_track *channels, size_t n, float *sum_squares, float *abs_max) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!channels || !sum_squares || !abs_max) return fail(ctx, THB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    std::lock_guard<std::mutex> lk(ctx->mu);
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

int thb_waveform_level(thb_ctx *ctx, const float *pcm, uint64_t len, uint64_t revision, uint32_t level, uint8_t *out,
                       size_t cap, size_t *written) {
    thb_track t{};
    t.pcm = pcm;
    t.len = len;
    uint8_t *outs[1] = {out};
    size_t caps[1] = {cap};
    size_t wr[1] = {0};
    if (!out) {
        if (written) *written = level_bytes(len, level);
        return THB_OK;
    }
    int rc = thb_waveform_level_batch(ctx, &t, 1, revision, level, outs, caps, wr, nullptr);
    if (written) *written = wr[0] ? wr[0] : level_bytes(len, level);
    return rc;
}

int thb_waveform_tile(thb_ctx *ctx, const float *pcm, uint64_t len, uint64_t revision, uint32_t level,
                      uint32_t tile_index, uint8_t *out, size_t cap, size_t *written) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    // header arithmetic of encode_waveform_tile with its saturating ops (render_tiles.rs:233-242)
    const uint64_t spb = level < 64 ? (1ull << level) : UINT64_MAX;
    const uint64_t tile_samples = spb > UINT64_MAX / 1024 ? UINT64_MAX : spb * 1024;
    const uint64_t start = (tile_index && tile_samples > UINT64_MAX / tile_index) ? UINT64_MAX : tile_samples * tile_index;
    const uint64_t end_unclamped = start > UINT64_MAX - tile_samples ? UINT64_MAX : start + tile_samples;
    const uint64_t end = std::min<uint64_t>(len, end_unclamped);
    const uint64_t bin_count = start >= end ? 0 : (end - start + spb - 1) / spb;
    const size_t need = 24 + 12 * bin_count;
    if (written) *written = need;
    if (!out) return THB_OK;
    if (cap < need) return fail(ctx, THB_ERR_SMALL_BUFFER, "need %zu bytes", need);
    if (bin_count == 0) {
        for (int i = 0; i < 8; i++) out[i] = static_cast<uint8_t>(revision >> (8 * i));
        put_u32(out + 8, 0);
        put_u32(out + 12, static_cast<uint32_t>(std::min<uint64_t>(spb, 0xffffffffull)));
        put_u32(out + 16, tile_index);
        put_u32(out + 20, 0);
        return THB_OK;
    }
    if (!pcm) return fail(ctx, THB_ERR_INVALID, "pcm is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    int rc = arena_begin(ctx, sizeof(thb::EnvDesc) + 1024);
    if (rc) return rc;
    thb::EnvDesc *d_desc = nullptr;
    thb::EnvDesc *h = arena_push<thb::EnvDesc>(ctx, 1, &d_desc);
    void *staging = nullptr;
    const float *d_pcm = pcm;
    if (!is_device_ptr(pcm)) {
        // only the tile's own samples cross PCIe; the kernel indexes from the file start
        const uint64_t cnt = end - start;
        CK(cudaMallocAsync(&staging, sizeof(float) * cnt + 64, ctx->stream));
        CK(cudaMemcpyAsync(staging, pcm + start, sizeof(float) * cnt, cudaMemcpyHostToDevice, ctx->stream));
        d_pcm = static_cast<const float *>(staging) - start;
    }
    void *d_out = nullptr;
    CK(cudaMallocAsync(&d_out, need, ctx->stream));
    h->pcm = d_pcm;
    h->len = static_cast<long long>(len);
    h->out = static_cast<uint8_t *>(d_out);
    if ((rc = arena_commit(ctx))) return rc;
    {
        ProfScope ps(ctx, "envelope");
        cudaError_t e = thb::launch_envelope(d_desc, 1, static_cast<long long>(len), level, revision, tile_index, 1, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "envelope: %s", cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(out, d_out, need, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaFreeAsync(d_out, ctx->stream));
    if (staging) CK(cudaFreeAsync(staging, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
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
    std::lock_guard<std::mutex> lk(ctx->mu);
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
    return THB_OK;
}

int thb_comm_destroy(thb_ctx *ctx) {
    if (!ctx) return THB_OK;
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
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->profiling = on != 0;
    return THB_OK;
}
int thb_profile_reset(thb_ctx *ctx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    ctx->prof.clear();
    ctx->launch_count = 0;
    return THB_OK;
}
int thb_profile_get(thb_ctx *ctx, const char *kernel, double *total_ms, uint64_t *launches) {
    if (!ctx || !kernel) return fail(ctx, THB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    auto it = ctx->prof.find(kernel);
    if (total_ms) *total_ms = it == ctx->prof.end() ? 0.0 : it->second.total_ms;
    if (launches) *launches = it == ctx->prof.end() ? 0 : it->second.launches;
    return THB_OK;
}
uint64_t thb_launch_count(const thb_ctx *ctx) { return ctx ? ctx->launch_count : 0; }

int thb_synth_pcm(thb_ctx *ctx, float *dev_out, uint64_t len, uint32_t sr, uint32_t track, uint32_t channel,
                  uint32_t flags) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (!dev_out || sr == 0 || len >= (1ull << 32)) return fail(ctx, THB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (!is_device_ptr(dev_out)) return fail(ctx, THB_ERR_INVALID, "dev_out must be device memory");
    cudaError_t e = thb::launch_synth_pcm(dev_out, len, sr, track, channel, flags, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "synth: %s", cudaGetErrorString(e));
    return THB_OK;
}

}  // extern "C"
