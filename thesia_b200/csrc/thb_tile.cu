// thb_tile.cu -- spectrogram tile encode: encode_spectrogram_tile / resize_spectrogram_tile
// (render_tiles.rs:281-393).  A tile is a <= 520 x 520 window of one level of detail of a retained u16 image,
// resampled with the separable Lanczos3 convolution the reference delegates to fast_image_resize (U16 pixels: i32
// fixed-point weights, i64 accumulation, half-LSB seed, arithmetic shift, clamp), mapped through the RGBA colormap
// and written with its rows reversed (high frequencies first).
//
//   pass H  tmp[y][x] = clip((2^(p-1) + sum_i img[y_first + y][start_x[x] + i] * wx[i][x]) >> p)   y over the rows pass V needs
//   pass V  px = clip((2^(q-1) + sum_i tmp[start_y[y] - y_first + i][x] * wy[i][y]) >> q) -> colormap -> out[height-1-y][x]
//
// Weights are tap-major (w[i][pixel]): in pass H neighbouring threads read neighbouring words, in pass V the weight
// is uniform over the CTA.  Integer arithmetic only: a tile is bit-identical to the CPU restatement.
#include "thb_device.cuh"
#include "thb_kernels.cuh"

#include <cstdlib>
#include <cstring>

namespace thb {
namespace {

constexpr int kTileThreads = 128;

__device__ __forceinline__ unsigned clip_u16(long long ss, unsigned precision) {
    const long long v = ss >> precision;
    return static_cast<unsigned>(v < 0 ? 0 : (v > 65535 ? 65535 : v));
}

// kTileRows rows per CTA: the weights of a column (pass H) are loaded once for all of them, and the descriptor,
// bounds and grid overhead are amortised
constexpr int kTileRows = 8;

__global__ void __launch_bounds__(kTileThreads) tile_horiz_kernel(const TileDesc *__restrict__ descs) {
    const TileDesc &d = descs[blockIdx.z];
    if (d.identity) return;  // tile_identity_kernel's
    const unsigned width = d.width, tmp_h = d.tmp_h;
    const unsigned x = blockIdx.x * kTileThreads + threadIdx.x;
    const unsigned y0 = blockIdx.y * kTileRows;
    if (y0 >= tmp_h || x >= width) return;
    const unsigned rows = min(static_cast<unsigned>(kTileRows), tmp_h - y0);
    const size_t pitch = d.pitch;
    const uint16_t *src = d.img + static_cast<size_t>(d.y_first + y0) * pitch + d.x_start[x];
    const unsigned n = d.x_size[x];
    const int *w = d.wx + x;
    const unsigned prec = d.px;
    long long ss[kTileRows];
#pragma unroll
    for (int r = 0; r < kTileRows; r++) ss[r] = 1ll << (prec - 1);
    if (rows == kTileRows) {
        for (unsigned i = 0; i < n; i++) {
            const long long k = __ldg(w + static_cast<size_t>(i) * width);
#pragma unroll
            for (int r = 0; r < kTileRows; r++) ss[r] += static_cast<long long>(__ldg(src + r * pitch + i)) * k;
        }
    } else {
        for (unsigned i = 0; i < n; i++) {
            const long long k = __ldg(w + static_cast<size_t>(i) * width);
            for (unsigned r = 0; r < rows; r++) ss[r] += static_cast<long long>(__ldg(src + r * pitch + i)) * k;
        }
    }
    uint16_t *dst = d.tmp + static_cast<size_t>(y0) * width + x;
#pragma unroll
    for (int r = 0; r < kTileRows; r++)
        if (static_cast<unsigned>(r) < rows) dst[static_cast<size_t>(r) * width] = static_cast<uint16_t>(clip_u16(ss[r], prec));
}

__global__ void __launch_bounds__(kTileThreads) tile_vert_kernel(const TileDesc *__restrict__ descs, const uchar4 *__restrict__ colormap,
                                                                 unsigned colors) {
    const TileDesc &d = descs[blockIdx.z];
    if (d.identity) return;  // tile_identity_kernel's
    const unsigned width = d.width, height = d.height;
    const unsigned x = blockIdx.x * kTileThreads + threadIdx.x;
    const unsigned y0 = blockIdx.y * kTileRows;
    if (y0 >= height || x >= width) return;
    const unsigned rows = min(static_cast<unsigned>(kTileRows), height - y0);
    const unsigned prec = d.py, y_first = d.y_first;
    const uint16_t *tmp = d.tmp + x;
    uchar4 *out = reinterpret_cast<uchar4 *>(d.out);
    for (unsigned r = 0; r < rows; r++) {
        const unsigned y = y0 + r;
        const unsigned n = d.y_size[y];
        const uint16_t *col = tmp + static_cast<size_t>(d.y_start[y] - y_first) * width;
        const int *w = d.wy + y;
        long long ss = 1ll << (prec - 1);
        for (unsigned i = 0; i < n; i++)
            ss += static_cast<long long>(col[static_cast<size_t>(i) * width]) * __ldg(w + static_cast<size_t>(i) * height);
        const unsigned long long v = clip_u16(ss, prec);
        // render_tiles.rs:339-346: (value * (color_count - 1) + u16::MAX / 2) / u16::MAX
        const unsigned ci = colors <= 1 ? 0u : static_cast<unsigned>((v * (colors - 1) + 32767ull) / 65535ull);
        out[static_cast<size_t>(height - 1 - y) * width + x] = __ldg(colormap + ci);
    }
}

// Level 0 in x and y: both resize passes are copies (one tap of weight 2^precision per output pixel: the fixed-point
// convolution returns the input value exactly), so the tile is colormap[index(img[y_first + y][x_first + x])] written
// with its rows reversed -- one pass, 2 bytes read and 4 written per pixel, no intermediate.  A thread takes four
// neighbouring pixels of kIdRows rows (8-byte loads and 16-byte stores when the addresses allow, which they do for
// the 512-pixel tile grid with its 4-pixel gutter); a CTA covers kIdRows whole rows of one tile.
constexpr int kIdRows = 8, kIdThreads = 160;   // 160 threads x 4 pixels cover a 520-pixel tile row in one pass
constexpr unsigned kIdSharedColors = 1024;     // SMEM variant: the colormap (the reference's has 258 entries) sits in shared memory
// SMEM: the lookups are the kernel's L1 traffic -- 32 lanes gathering from a 1 KB table are up to nine 128-byte lines per
// instruction through the global path, against ~3.5 bank-conflict wavefronts from shared memory.  A CTA then takes
// `groups` row groups of one tile so that filling the table (one L2 read of the colormap per CTA) is amortised.
template <bool SMEM>
__global__ void __launch_bounds__(kIdThreads) tile_identity_kernel(const TileDesc *__restrict__ descs, const uchar4 *__restrict__ colormap,
                                                                   unsigned colors, unsigned groups) {
    __shared__ unsigned s_cm[SMEM ? kIdSharedColors : 1];
    const TileDesc &d = descs[blockIdx.y];
    if (!d.identity) return;
    const unsigned width = d.width, height = d.height;
    if (blockIdx.x * kIdRows * groups >= height) return;
    const unsigned *cm = reinterpret_cast<const unsigned *>(colormap);   // about 1 KB
    if (SMEM) {
        for (unsigned i = threadIdx.x; i < colors; i += kIdThreads) s_cm[i] = __ldg(cm + i);
        __syncthreads();
    }
    const size_t pitch = d.pitch;
    unsigned *out = reinterpret_cast<unsigned *>(d.out);
    const unsigned scale = colors - 1;
    const bool small_map = colors <= 65536u;
    auto fetch = [&](unsigned ci) -> unsigned { return SMEM ? s_cm[ci] : __ldg(cm + ci); };
    auto look = [&](unsigned v) -> unsigned {
        // render_tiles.rs:339-346: (value * (color_count - 1) + u16::MAX / 2) / u16::MAX  (32 bits hold it when colors <= 2^16)
        const unsigned ci = colors <= 1 ? 0u : (small_map ? (v * scale + 32767u) / 65535u
                                                          : static_cast<unsigned>((static_cast<unsigned long long>(v) * scale + 32767ull) / 65535ull));
        return fetch(ci);
    };
    // (straight-line 32-bit index arithmetic for the common case: a branch per lookup puts every load in its own basic block)
    auto look32 = [&](unsigned v) -> unsigned { return fetch((v * scale + 32767u) / 65535u); };
    const bool out_aligned = (width & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (unsigned g = 0; g < groups; g++) {
        const unsigned y0 = (blockIdx.x * groups + g) * kIdRows;
        if (y0 >= height) break;
        const unsigned rows = min(static_cast<unsigned>(kIdRows), height - y0);
        const uint16_t *src0 = d.img + static_cast<size_t>(d.y_first + y0) * pitch + d.x_first;
        // rows of the image are 8-byte aligned relative to each other when the pitch is a multiple of four pixels
        const bool rows_aligned = (pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(src0) & 7) == 0;
        for (unsigned x = 4 * threadIdx.x; x < width; x += 4 * kIdThreads) {
            const unsigned npx = min(4u, width - x);
            if (npx == 4 && rows_aligned && rows == kIdRows && colors > 1 && small_map) {
                // the common case: the eight 8-byte loads of the thread, the 32 lookups, the eight 16-byte stores, with
                // no control flow in between (the stores may alias the loads as far as the compiler knows: a per-row
                // loop serialises load -> lookup -> store eight times)
                uint2 w[kIdRows];
#pragma unroll
                for (int r = 0; r < kIdRows; r++) w[r] = __ldg(reinterpret_cast<const uint2 *>(src0 + static_cast<size_t>(r) * pitch + x));
                uint4 c[kIdRows];
#pragma unroll
                for (int r = 0; r < kIdRows; r++)
                    c[r] = make_uint4(look32(w[r].x & 0xffffu), look32(w[r].x >> 16), look32(w[r].y & 0xffffu), look32(w[r].y >> 16));
#pragma unroll
                for (int r = 0; r < kIdRows; r++) {
                    unsigned *o = out + static_cast<size_t>(height - 1 - (y0 + r)) * width + x;
                    if (out_aligned) {
                        *reinterpret_cast<uint4 *>(o) = c[r];
                    } else {
                        o[0] = c[r].x; o[1] = c[r].y; o[2] = c[r].z; o[3] = c[r].w;
                    }
                }
                continue;
            }
            for (unsigned r = 0; r < rows; r++) {
                const uint16_t *s = src0 + static_cast<size_t>(r) * pitch + x;
                unsigned *o = out + static_cast<size_t>(height - 1 - (y0 + r)) * width + x;
                if (npx == 4 && (reinterpret_cast<uintptr_t>(s) & 7) == 0) {
                    const uint2 w = __ldg(reinterpret_cast<const uint2 *>(s));
                    const uint4 c = make_uint4(look(w.x & 0xffffu), look(w.x >> 16), look(w.y & 0xffffu), look(w.y >> 16));
                    if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                        *reinterpret_cast<uint4 *>(o) = c;
                    } else {
                        o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
                    }
                } else {
                    for (unsigned i = 0; i < npx; i++) o[i] = look(__ldg(s + i));
                }
            }
        }
    }
}

}  // namespace

int spectrogram_tile_launches(int n, int n_identity) {
    const int chunks = (n + 65534) / 65535;
    return chunks * ((n_identity > 0 ? 1 : 0) + (n_identity < n ? 2 : 0));
}

cudaError_t launch_spectrogram_tiles(const TileDesc *d_descs, int n, int n_identity, unsigned max_w, unsigned max_h, unsigned max_tmp_h,
                                     const uchar4 *d_colormap, unsigned colors, cudaStream_t st) {
    if (n <= 0 || !max_w || !max_h) return cudaSuccess;
    const unsigned gx = (max_w + kTileThreads - 1) / kTileThreads;
    for (int c0 = 0; c0 < n; c0 += 65535) {
        const unsigned nc = static_cast<unsigned>(n - c0 < 65535 ? n - c0 : 65535);
        // each kernel returns at once for the descriptors of the other kind -- but a grid of CTAs that only return is
        // not free (a level-0 batch of 1 760 tiles made the two resample kernels walk 2 x 396 000 of them, a third of
        // the batch's time), so a kind that does not occur in the batch is not launched at all
        cudaError_t e = cudaSuccess;
        if (n_identity > 0) {
            // THB_TILE_CM = l1 | smem (default smem when the colormap fits), THB_TILE_GROUPS = row groups per CTA
            static const char *cm_env = getenv("THB_TILE_CM");
            static const char *gr_env = getenv("THB_TILE_GROUPS");
            const bool smem = colors <= kIdSharedColors && !(cm_env && !strcmp(cm_env, "l1"));
            unsigned groups = smem ? 4u : 1u;
            if (gr_env && atoi(gr_env) > 0) groups = static_cast<unsigned>(atoi(gr_env));
            const dim3 grid((max_h + kIdRows * groups - 1) / (kIdRows * groups), nc);
            if (smem) tile_identity_kernel<true><<<grid, kIdThreads, 0, st>>>(d_descs + c0, d_colormap, colors, groups);
            else tile_identity_kernel<false><<<grid, kIdThreads, 0, st>>>(d_descs + c0, d_colormap, colors, groups);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        if (n_identity >= n) continue;
        tile_horiz_kernel<<<dim3(gx, (max_tmp_h + kTileRows - 1) / kTileRows, nc), kTileThreads, 0, st>>>(d_descs + c0);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        tile_vert_kernel<<<dim3(gx, (max_h + kTileRows - 1) / kTileRows, nc), kTileThreads, 0, st>>>(d_descs + c0, d_colormap, colors);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace thb
