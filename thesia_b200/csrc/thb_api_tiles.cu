// thb_api_tiles.cu -- the tile readers of the C ABI (include/thesia_b200.h): encode_waveform_tile and
// encode_spectrogram_tile (render_tiles.rs:232-393).
//
// The reference serves tiles from concurrent IPC threads under READ locks while the analysis runs under the write lock
// (lib.rs:343-389, interface.rs:12-56).  Same here: thb_waveform_tile / thb_spectrogram_tile[_batch] take ctx->mu shared
// and each call works on its own LANE -- a stream with its own pinned / device descriptor buffer and scratch -- so
// concurrent calls overlap on the device and never wait behind a thb_spec_batch that is not running.  Host PCM a tile
// call has seen stays on the device (PcmCache, keyed by (pointer, length, revision), filled granule by granule), so a
// second tile of the same channel does not cross PCIe again.
#include "thb_ctx.hpp"

namespace thbapi {

struct TileLane {
    cudaStream_t stream = nullptr;
    unsigned char *h_buf = nullptr, *d_buf = nullptr;  // descriptors, colormap, small outputs
    size_t buf_cap = 0;
    unsigned char *d_scratch = nullptr;                // staged PCM / resize intermediates / RGBA tiles
    size_t scratch_cap = 0;
    bool busy = false;
    TileLane() = default;
    TileLane(const TileLane &) = delete;
    TileLane &operator=(const TileLane &) = delete;
    ~TileLane() {
        if (stream) cudaStreamSynchronize(stream);
        if (h_buf) cudaFreeHost(h_buf);
        if (d_buf) cudaFree(d_buf);
        if (d_scratch) cudaFree(d_scratch);
        if (stream) cudaStreamDestroy(stream);
    }
    cudaError_t reserve_buf(size_t need) {
        if (need <= buf_cap) return cudaSuccess;
        cudaStreamSynchronize(stream);
        if (h_buf) cudaFreeHost(h_buf);
        if (d_buf) cudaFree(d_buf);
        h_buf = d_buf = nullptr;
        buf_cap = 0;
        size_t cap = 1 << 16;
        while (cap < need) cap <<= 1;
        cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&h_buf), cap);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&d_buf), cap);
        if (e == cudaSuccess) buf_cap = cap;
        return e;
    }
    cudaError_t reserve_scratch(size_t need) {
        if (need <= scratch_cap) return cudaSuccess;
        cudaStreamSynchronize(stream);
        if (d_scratch) cudaFree(d_scratch);
        d_scratch = nullptr;
        scratch_cap = 0;
        size_t cap = 1 << 20;
        while (cap < need) cap <<= 1;
        const cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&d_scratch), cap);
        if (e == cudaSuccess) scratch_cap = cap;
        return e;
    }
};

// one lane per call in flight; released when the call returns
struct LaneLease {
    thb_ctx *ctx = nullptr;
    TileLane *lane = nullptr;
    LaneLease() = default;
    LaneLease(const LaneLease &) = delete;
    LaneLease &operator=(const LaneLease &) = delete;
    ~LaneLease() {
        if (!lane) return;
        std::lock_guard<std::mutex> lk(ctx->lane_mu);
        lane->busy = false;
    }
};

static int lane_acquire(thb_ctx *ctx, LaneLease *out) {
    std::lock_guard<std::mutex> lk(ctx->lane_mu);
    for (auto &l : ctx->lanes)
        if (!l->busy) {
            l->busy = true;
            out->ctx = ctx;
            out->lane = l.get();
            return THB_OK;
        }
    auto l = std::make_unique<TileLane>();
    CK(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
    CK(l->reserve_buf(1 << 16));
    l->busy = true;
    out->ctx = ctx;
    out->lane = l.get();
    ctx->lanes.push_back(std::move(l));
    return THB_OK;
}

// Device copies of host channels the tile readers have been handed, keyed by (host pointer, length, revision) -- the
// key of the reference's own tile cache (render_tiles.rs:124-169).  A channel's buffer is allocated once and filled
// granule by granule with exactly the samples tile calls touch; a new revision or length of the same pointer drops
// the old copy.  Bounded by THB_PCM_CACHE_MB (default 4096, 0 disables), least recently used first.
struct PcmEntry {
    const void *host = nullptr;
    uint64_t len = 0, revision = 0;
    float *d_pcm = nullptr;
    size_t bytes = 0;
    uint64_t last_use = 0;
    std::mutex mu;               // guards `valid` and the uploads that set it
    std::vector<uint8_t> valid;  // per granule
    ~PcmEntry() {
        if (d_pcm) cudaFree(d_pcm);
    }
};
constexpr uint64_t kGranule = 1ull << 16;  // samples (256 KB)

struct PcmCache {
    std::mutex mu;
    std::list<std::shared_ptr<PcmEntry>> entries;
    size_t bytes = 0, cap_bytes = 0;
    uint64_t tick = 0;
    uint64_t hits = 0, misses = 0;
};

static std::shared_ptr<PcmEntry> pcm_lookup(thb_ctx *ctx, const float *pcm, uint64_t len, uint64_t revision) {
    PcmCache &c = *ctx->pcm_cache;
    const size_t bytes = sizeof(float) * len;
    if (c.cap_bytes == 0 || bytes > c.cap_bytes) return nullptr;
    std::lock_guard<std::mutex> lk(c.mu);
    c.tick++;
    for (auto it = c.entries.begin(); it != c.entries.end();) {
        PcmEntry &e = **it;
        if (e.host == pcm) {
            if (e.len == len && e.revision == revision) {
                e.last_use = c.tick;
                return *it;
            }
            c.bytes -= e.bytes;  // the channel changed: its old copy is dead (freed when the last reader lets go)
            it = c.entries.erase(it);
            continue;
        }
        ++it;
    }
    while (c.bytes + bytes > c.cap_bytes && !c.entries.empty()) {
        auto lru = c.entries.begin();
        for (auto it = c.entries.begin(); it != c.entries.end(); ++it)
            if ((*it)->last_use < (*lru)->last_use) lru = it;
        c.bytes -= (*lru)->bytes;
        c.entries.erase(lru);
    }
    auto e = std::make_shared<PcmEntry>();
    if (cudaMalloc(reinterpret_cast<void **>(&e->d_pcm), bytes ? bytes : 4) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;  // no room: the call stages its own samples instead
    }
    e->host = pcm;
    e->len = len;
    e->revision = revision;
    e->bytes = bytes;
    e->last_use = c.tick;
    e->valid.assign(static_cast<size_t>((len + kGranule - 1) / kGranule), 0);
    c.bytes += bytes;
    c.entries.push_back(e);
    return e;
}

// samples [start, end) of the entry's channel are on the device when this returns
static int pcm_ensure(thb_ctx *ctx, PcmEntry &e, const float *pcm, uint64_t start, uint64_t end, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(e.mu);
    const size_t g0 = static_cast<size_t>(start / kGranule), g1 = static_cast<size_t>((end + kGranule - 1) / kGranule);
    bool any = false;
    for (size_t g = g0; g < g1;) {
        if (e.valid[g]) {
            g++;
            continue;
        }
        size_t h = g;
        while (h < g1 && !e.valid[h]) h++;
        const uint64_t s0 = g * kGranule, s1 = std::min<uint64_t>(e.len, h * kGranule);
        CK(cudaMemcpyAsync(e.d_pcm + s0, pcm + s0, sizeof(float) * (s1 - s0), cudaMemcpyHostToDevice, st));
        any = true;
        g = h;
    }
    {
        std::lock_guard<std::mutex> ck(ctx->pcm_cache->mu);
        (any ? ctx->pcm_cache->misses : ctx->pcm_cache->hits)++;
    }
    if (!any) return THB_OK;
    CK(cudaStreamSynchronize(st));  // the host samples belong to the caller again; other lanes may now read the copy
    for (size_t g = g0; g < g1; g++) e.valid[g] = 1;
    return THB_OK;
}

void tiles_shutdown(thb_ctx *ctx) {
    {
        std::lock_guard<std::mutex> lk(ctx->lane_mu);
        ctx->lanes.clear();
    }
    if (ctx->pcm_cache) {
        std::lock_guard<std::mutex> lk(ctx->pcm_cache->mu);
        ctx->pcm_cache->entries.clear();
        ctx->pcm_cache->bytes = 0;
    }
}

static int get_tile_axis(thb_ctx *ctx, cudaStream_t st, uint32_t in_size, uint64_t lod_size, uint64_t origin, uint32_t out_size,
                         std::shared_ptr<thb_ctx::AxisDev> *out) {
    const auto key = std::make_tuple(in_size, lod_size, origin, out_size);
    {
        std::lock_guard<std::mutex> lk(ctx->axes_mu);
        auto it = ctx->tile_axes.find(key);
        if (it != ctx->tile_axes.end()) {
            *out = it->second;
            return THB_OK;
        }
    }
    // render_tiles.rs:379-383: the crop box in source pixels, f64
    const double in0 = static_cast<double>(origin) * static_cast<double>(in_size) / static_cast<double>(lod_size);
    const double in1 = static_cast<double>(origin + out_size) * static_cast<double>(in_size) / static_cast<double>(lod_size);
    const thb::ResizeAxis a = thb::resize_axis(in_size, in0, in0 + (in1 - in0), out_size);
    auto d = std::make_shared<thb_ctx::AxisDev>();
    d->n = a.n;
    d->window = a.window;
    d->precision = a.precision;
    d->first = UINT32_MAX;
    d->end = 0;
    d->identity = a.n > 0;
    for (uint32_t o = 0; o < a.n; o++) {
        d->first = std::min(d->first, a.start[o]);
        d->end = std::max(d->end, a.start[o] + a.size[o]);
        if (a.size[o] != 1 || a.start[o] != a.start[0] + o || a.w_t[o] != (int32_t(1) << a.precision)) d->identity = false;
    }
    CK(cudaMalloc(reinterpret_cast<void **>(&d->start), sizeof(unsigned) * a.n));
    CK(cudaMalloc(reinterpret_cast<void **>(&d->size), sizeof(unsigned) * a.n));
    CK(cudaMalloc(reinterpret_cast<void **>(&d->w), sizeof(int) * a.w_t.size()));
    // the vectors die with this scope: plain (staged) copies, complete before the entry becomes visible
    CK(cudaMemcpyAsync(d->start, a.start.data(), sizeof(unsigned) * a.n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->size, a.size.data(), sizeof(unsigned) * a.n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->w, a.w_t.data(), sizeof(int) * a.w_t.size(), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    std::lock_guard<std::mutex> lk(ctx->axes_mu);
    if (ctx->tile_axes.size() >= 1024) ctx->tile_axes.clear();  // bound the cache; readers keep their own references
    auto ins = ctx->tile_axes.emplace(key, d);
    *out = ins.first->second;
    return THB_OK;
}

}  // namespace thbapi

thb_ctx::thb_ctx() : pcm_cache(new PcmCache()) {
    size_t mb = 4096;
    if (const char *e = getenv("THB_PCM_CACHE_MB")) mb = static_cast<size_t>(std::max(0ll, atoll(e)));
    pcm_cache->cap_bytes = mb << 20;
}
thb_ctx::~thb_ctx() = default;

extern "C" {

// ---- spectrogram tiles (render_tiles.rs:281-393) ----------------------------------------------------
int thb_spectrogram_tile_geometry(uint64_t height, uint64_t width, uint32_t level_x, uint32_t level_y, uint32_t tile_x,
                                  uint32_t tile_y, uint64_t geo[6]) {
    if (!geo) return fail(nullptr, THB_ERR_INVALID, "geo is NULL");
    const thb::TileGeometry g = thb::spectrogram_tile_geometry(height, width, level_x, level_y, tile_x, tile_y);
    geo[0] = g.lod_width; geo[1] = g.lod_height; geo[2] = g.origin_x; geo[3] = g.origin_y; geo[4] = g.width; geo[5] = g.height;
    return THB_OK;
}

int thb_spectrogram_tile_batch(thb_ctx *ctx, const uint8_t *colormap_rgba, size_t colormap_bytes, uint64_t revision,
                               thb_spec_tile_req *reqs, size_t n) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    if (n == 0) return THB_OK;
    if (!reqs) return fail(ctx, THB_ERR_INVALID, "reqs is NULL");
    if (!colormap_rgba || colormap_bytes < 4 || colormap_bytes % 4)  // RenderTileCache::set_colormap (render_tiles.rs:80-85)
        return fail(ctx, THB_ERR_INVALID, "colormap must be a non-empty RGBA table");
    Nvtx nv("thb_spectrogram_tile_batch");
    ReadLock lk(ctx->mu);  // the images cannot change under us; other tile readers run alongside
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
    LaneLease lease;
    int rc = lane_acquire(ctx, &lease);
    if (rc) return rc;
    TileLane &lane = *lease.lane;
    const size_t m = work.size();
    std::vector<std::shared_ptr<thb_ctx::AxisDev>> axs(m), ays(m);
    for (size_t k = 0; k < m; k++) {
        const Work &w = work[k];
        rc = get_tile_axis(ctx, lane.stream, static_cast<uint32_t>(w.sp->T), w.g.lod_width, w.g.origin_x, static_cast<uint32_t>(w.g.width), &axs[k]);
        if (!rc) rc = get_tile_axis(ctx, lane.stream, static_cast<uint32_t>(w.sp->img_H), w.g.lod_height, w.g.origin_y, static_cast<uint32_t>(w.g.height), &ays[k]);
        if (rc) return rc;
    }
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
    // lane buffer: [descriptors | colormap]; lane scratch: [tmp (u16) | out (RGBA)]
    const size_t desc_bytes = (sizeof(thb::TileDesc) * m + 255) & ~size_t(255);
    const size_t tmp_bytes = (sizeof(uint16_t) * tmp_total + 16 + 255) & ~size_t(255);
    CK(lane.reserve_buf(desc_bytes + colormap_bytes));
    CK(lane.reserve_scratch(tmp_bytes + out_total + 16));
    thb::TileDesc *h = reinterpret_cast<thb::TileDesc *>(lane.h_buf);
    thb::TileDesc *d_desc = reinterpret_cast<thb::TileDesc *>(lane.d_buf);
    memcpy(lane.h_buf + desc_bytes, colormap_rgba, colormap_bytes);
    const uchar4 *d_cm = reinterpret_cast<const uchar4 *>(lane.d_buf + desc_bytes);
    uint16_t *d_tmp = reinterpret_cast<uint16_t *>(lane.d_scratch);
    uint8_t *d_out = lane.d_scratch + tmp_bytes;
    size_t tmp_off = 0, out_off = 0;
    int n_identity = 0;
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
        t.identity = axs[k]->identity && ays[k]->identity ? 1u : 0u;
        n_identity += static_cast<int>(t.identity);
        t.x_first = axs[k]->first;
        tmp_off += (static_cast<size_t>(t.tmp_h) * t.width + 7) & ~size_t(7);
        out_off += static_cast<size_t>(t.width) * t.height * 4;
    }
    CK(cudaMemcpyAsync(lane.d_buf, lane.h_buf, desc_bytes + colormap_bytes, cudaMemcpyHostToDevice, lane.stream));
    {
        ProfScope ps(ctx, "spectrogram_tile", thb::spectrogram_tile_launches(static_cast<int>(m), n_identity), lane.stream);
        cudaError_t e = thb::launch_spectrogram_tiles(d_desc, static_cast<int>(m), n_identity, max_w, max_h, max_tmp_h, d_cm,
                                                      static_cast<unsigned>(colormap_bytes / 4), lane.stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "spectrogram_tile: %s", cudaGetErrorString(e));
    }
    out_off = 0;
    for (size_t k = 0; k < m; k++) {
        const size_t bytes = static_cast<size_t>(work[k].g.width) * work[k].g.height * 4;
        CK(cudaMemcpyAsync(reqs[work[k].req].out + 40, d_out + out_off, bytes, cudaMemcpyDeviceToHost, lane.stream));
        out_off += bytes;
    }
    CK(cudaStreamSynchronize(lane.stream));
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
    Nvtx nv("thb_waveform_level_batch");
    WriteLock lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    for (void *p : ctx->env_outputs) CK(cudaFreeAsync(p, ctx->stream));
    ctx->env_outputs.clear();
    int rc = arena_begin(ctx, (sizeof(thb::EnvDesc) + 64) * n + 1024);
    if (rc) return rc;
    thb::EnvDesc *d_desc = nullptr;
    thb::EnvDesc *h = arena_push<thb::EnvDesc>(ctx, n, &d_desc);
    Scratch staging(ctx);  // host channels are staged; the buffers go back to the pool when the call ends
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
            float *st = nullptr;
            CK(staging.alloc(&st, sizeof(float) * t.len + 64));
            CK(cudaMemcpyAsync(st, t.pcm, sizeof(float) * t.len, cudaMemcpyHostToDevice, ctx->stream));
            d_pcm = st;
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
        if (host_out && host_out[i] && bytes[i]) {
            CK(cudaMemcpyAsync(host_out[i], ctx->env_outputs[i], bytes[i], cudaMemcpyDeviceToHost, ctx->stream));
            any_out = true;
        }
    }
    if (any_out || any_host) CK(cudaStreamSynchronize(ctx->stream));
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
    Nvtx nv("thb_waveform_tile");
    ReadLock lk(ctx->mu);  // concurrent with other tile readers; only the analysis (write lock) excludes us
    CK(cudaSetDevice(ctx->device));
    LaneLease lease;
    int rc = lane_acquire(ctx, &lease);
    if (rc) return rc;
    TileLane &lane = *lease.lane;
    constexpr size_t kOutOff = 256;
    CK(lane.reserve_buf(kOutOff + need));
    const float *d_pcm = pcm;
    std::shared_ptr<PcmEntry> cached;
    if (!is_device_ptr(pcm)) {
        cached = pcm_lookup(ctx, pcm, len, revision);
        if (cached) {
            // only the granules this tile touches and the device has not seen yet cross PCIe
            if ((rc = pcm_ensure(ctx, *cached, pcm, start, end, lane.stream))) return rc;
            d_pcm = cached->d_pcm;
        } else {
            // cache off / channel too large: only the tile's own samples cross PCIe; the kernel indexes from the file start
            const uint64_t cnt = end - start;
            CK(lane.reserve_scratch(sizeof(float) * cnt + 64));
            CK(cudaMemcpyAsync(lane.d_scratch, pcm + start, sizeof(float) * cnt, cudaMemcpyHostToDevice, lane.stream));
            d_pcm = reinterpret_cast<const float *>(lane.d_scratch) - start;
        }
    }
    thb::EnvDesc *h = reinterpret_cast<thb::EnvDesc *>(lane.h_buf);
    h->pcm = d_pcm;
    h->len = static_cast<long long>(len);
    h->out = lane.d_buf + kOutOff;
    CK(cudaMemcpyAsync(lane.d_buf, lane.h_buf, sizeof(thb::EnvDesc), cudaMemcpyHostToDevice, lane.stream));
    {
        ProfScope ps(ctx, "envelope", 1, lane.stream);
        cudaError_t e = thb::launch_envelope(reinterpret_cast<const thb::EnvDesc *>(lane.d_buf), 1, static_cast<long long>(len), level,
                                             revision, tile_index, 1, lane.stream);
        if (e != cudaSuccess) return fail(ctx, THB_ERR_CUDA, "envelope: %s", cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(lane.h_buf + kOutOff, lane.d_buf + kOutOff, need, cudaMemcpyDeviceToHost, lane.stream));
    CK(cudaStreamSynchronize(lane.stream));
    memcpy(out, lane.h_buf + kOutOff, need);
    return THB_OK;
}

int thb_pcm_cache_stats(thb_ctx *ctx, uint64_t *entries, uint64_t *bytes, uint64_t *hits, uint64_t *misses) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->pcm_cache->mu);
    if (entries) *entries = ctx->pcm_cache->entries.size();
    if (bytes) *bytes = ctx->pcm_cache->bytes;
    if (hits) *hits = ctx->pcm_cache->hits;
    if (misses) *misses = ctx->pcm_cache->misses;
    return THB_OK;
}

int thb_pcm_cache_clear(thb_ctx *ctx) {
    if (!ctx) return fail(nullptr, THB_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->pcm_cache->mu);
    ctx->pcm_cache->entries.clear();
    ctx->pcm_cache->bytes = 0;
    return THB_OK;
}

}  // extern "C"
