// thb_ctx.hpp -- state and helpers shared by the files that implement the extern "C" surface of
// include/thesia_b200.h (thb_api_core.cu, thb_api_spec.cu, thb_api_tiles.cu, thb_api_dynamics.cu).
//
// Host-side orchestration only: plan cache (SpectrogramAnalyzer::prepare, spectrogram.rs:116-154),
// the (id, ch) -> spectrogram store (TrackManager.specs / spec_imgs, mod.rs:33-44), stream-ordered
// device memory, descriptor upload, kernel launches, the 2-float NCCL all-reduce and measurement
// hooks.  All arithmetic on samples happens in the kernels; there is no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
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

namespace thbapi {

struct PlanKey {
    uint32_t sr;
    uint64_t hop, win, n_fft;
    uint32_t freq_scale, n_mel_req;
    bool operator<(const PlanKey &o) const {
        return std::tie(sr, hop, win, n_fft, freq_scale, n_mel_req) <
               std::tie(o.sr, o.hop, o.win, o.n_fft, o.freq_scale, o.n_mel_req);
    }
};

// tables of one analyzer plan on the device; the allocations die with the plan, whatever the exit path
struct Plan {
    thb::PlanDev dev{};
    std::vector<void *> allocs;
    Plan() = default;
    Plan(const Plan &) = delete;
    Plan &operator=(const Plan &) = delete;
    ~Plan() {
        for (void *p : allocs) cudaFree(p);
    }
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
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string *err);
};
extern NcclApi g_nccl;
constexpr int kNcclFloat32 = 7;
constexpr int kNcclMax = 2;
constexpr int kNcclMin = 3;
constexpr int kNcclUint8 = 1;
constexpr int kNcclInt32 = 2;

struct TileLane;
struct PcmCache;

}  // namespace thbapi

using namespace thbapi;

struct thb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // H2D staging, overlapped with compute
    cudaStream_t side_stream = nullptr;  // the scalar kernel over file-edge frames, next to the packed kernel
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    // Writers (update_specs, update_spec_imgs, release, ...) take `mu` exclusively; the tile readers
    // (thb_waveform_tile, thb_spectrogram_tile[_batch]) take it shared and run concurrently on their own
    // streams -- the reference's write-lock worker / read-locked IPC threads (interface.rs:12-56, lib.rs:343-389).
    std::shared_mutex mu;
    std::mutex err_mu;
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

    // descriptor arenas: pinned host mirror + device copy (writers only).  A ring of four: a call fills the next one
    // while the uploads of the previous calls may still be queued behind their kernels, so a host that queues steps
    // back to back (thb_update_spec_imgs without a host round trip) is never held up by its own previous step.
    static constexpr int kArenas = 4;
    struct Arena {
        unsigned char *h = nullptr, *d = nullptr;
        size_t cap = 0;
        cudaEvent_t ev = nullptr;  // recorded after the upload: the pinned mirror may be rewritten once it has passed
    };
    Arena arenas[kArenas];
    int arena_idx = 0;
    unsigned char *h_arena = nullptr, *d_arena = nullptr;  // the current one
    size_t arena_used = 0;
    cudaEvent_t h2d_ev = nullptr;
    std::vector<cudaEvent_t> stage_ev;  // one per H2D pipeline stage of thb_spec_batch

    // tiles the frame-pair STFT kernel hands back to the scalar kernel (thb_kernels.cuh RescueList)
    uint2 *d_rescue_items = nullptr;
    unsigned *d_rescue_count = nullptr;  // [0] = count, [1 ..] = one flag per (descriptor, tile)
    size_t rescue_cap = 0;

    // the quantiser's descriptors of the last thb_update_spec_imgs: re-quantising the same retained spectrograms (a new
    // dB range, another colormap length, the next step of a batch loop) reuses the device copy instead of uploading it
    // between the range exchange and the quantiser
    std::vector<unsigned char> img_desc_host;
    unsigned char *d_img_desc = nullptr;
    size_t img_desc_cap = 0;

    std::vector<void *> env_outputs;  // device buffers of the last waveform level call

    // resize axes of the spectrogram tiles (thb_host.hpp ResizeAxis) on the device, keyed by
    // (in_size, lod_size, origin, out_size): every tile of one tile row / column of a level shares one.
    // Entries are immutable once built and reference counted: a reader keeps its axes alive while the cache is trimmed.
    struct AxisDev {
        unsigned *start = nullptr, *size = nullptr;
        int *w = nullptr;
        unsigned n = 0, window = 0, precision = 0, first = 0, end = 0;
        bool identity = false;  // every output pixel is one input pixel with weight 2^precision, starts consecutive
        AxisDev() = default;
        AxisDev(const AxisDev &) = delete;
        AxisDev &operator=(const AxisDev &) = delete;
        ~AxisDev() {
            cudaFree(start);
            cudaFree(size);
            cudaFree(w);
        }
    };
    std::mutex axes_mu;
    std::map<std::tuple<uint32_t, uint64_t, uint64_t, uint32_t>, std::shared_ptr<AxisDev>> tile_axes;

    // tile readers: one lane (stream + descriptor / scratch buffers) per concurrent call, and the device copies of
    // the PCM they have seen (thb_api_tiles.cu)
    std::mutex lane_mu;
    std::vector<std::unique_ptr<TileLane>> lanes;
    std::unique_ptr<PcmCache> pcm_cache;

    void *nccl_comm = nullptr;
    int n_ranks = 1, rank = 0;
    // the global-range exchange over peer memory (thb_comm_init): every rank's slots mapped through CUDA IPC
    struct PeerExchange {
        bool ok = false;
        float4 *mine = nullptr;                 // [2][kExchangeMaxRanks] {max, -min, seq, -}
        float4 **d_peers = nullptr;             // device array of n_ranks pointers (peers[rank] == mine)
        std::vector<void *> opened;             // IPC mappings to close
        unsigned *d_fail = nullptr;
        unsigned seq = 0;
    } px;

    std::mutex prof_mu;
    bool profiling = false;
    std::map<std::string, ProfEntry> prof;
    std::vector<cudaEvent_t> event_pool;
    std::atomic<uint64_t> launch_count{0};

    thb_ctx();
    ~thb_ctx();
};

namespace thbapi {

extern thread_local std::string g_last_error;
int fail(thb_ctx *ctx, int code, const char *fmt, ...) __attribute__((format(printf, 3, 4)));

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? THB_ERR_NOMEM : THB_ERR_CUDA,        \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

using WriteLock = std::unique_lock<std::shared_mutex>;
using ReadLock = std::shared_lock<std::shared_mutex>;

// NVTX range around every C-ABI entry point (a no-op unless a profiler is attached)
struct Nvtx {
    explicit Nvtx(const char *name) { nvtxRangePushA(name); }
    Nvtx(const Nvtx &) = delete;
    Nvtx &operator=(const Nvtx &) = delete;
    ~Nvtx() { nvtxRangePop(); }
};

bool is_device_ptr(const void *p);

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

// ---- measurement hooks: CUDA events around a launch group on the stream it is launched on ----
struct ProfScope {
    thb_ctx *ctx;
    ProfEntry *e = nullptr;
    cudaStream_t st;
    cudaEvent_t start = nullptr, stop = nullptr;
    ProfScope(thb_ctx *c, const char *name, int launches = 1, cudaStream_t stream = nullptr);
    ProfScope(const ProfScope &) = delete;
    ProfScope &operator=(const ProfScope &) = delete;
    ~ProfScope();
};
void prof_collect(thb_ctx *ctx);  // caller holds prof_mu

// ---- descriptor arena (writers only) ----
int arena_begin(thb_ctx *ctx, size_t need);
template <typename T>
T *arena_push(thb_ctx *ctx, size_t count, T **dev) {
    size_t off = (ctx->arena_used + 255) & ~size_t(255);
    ctx->arena_used = off + sizeof(T) * count;
    *dev = reinterpret_cast<T *>(ctx->d_arena + off);
    return reinterpret_cast<T *>(ctx->h_arena + off);
}
int arena_commit(thb_ctx *ctx);

int get_plan(thb_ctx *ctx, const thb_setting &s, uint32_t sr, const Plan **out);
int slot_alloc(thb_ctx *ctx, int *slot);
void spec_free(thb_ctx *ctx, Spec &s);
Spec *find_spec(thb_ctx *ctx, uint64_t id, uint32_t ch);
int global_minmax_on_stream(thb_ctx *ctx, float dB_range);
uint64_t level_bytes(uint64_t len, uint32_t level);
void put_u32(uint8_t *p, uint32_t v);
void tiles_shutdown(thb_ctx *ctx);  // thb_api_tiles.cu: lanes + PCM cache

}  // namespace thbapi
