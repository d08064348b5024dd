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
__global__ void __launch_bounds__(kIdThreads) tile_identity_kernel(const TileDesc *__restrict__ descs, const uchar4 *__restrict__ colormap,
                                                                   unsigned colors) {
    const TileDesc &d = descs[blockIdx.y];
    if (!d.identity) return;
    const unsigned width = d.width, height = d.height;
    const unsigned y0 = blockIdx.x * kIdRows;
    if (y0 >= height) return;
    const unsigned rows = min(static_cast<unsigned>(kIdRows), height - y0);
    const size_t pitch = d.pitch;
    const uint16_t *src0 = d.img + static_cast<size_t>(d.y_first + y0) * pitch + d.x_first;
    unsigned *out = reinterpret_cast<unsigned *>(d.out);
    const unsigned *cm = reinterpret_cast<const unsigned *>(colormap);   // about 1 KB, read through L1
    const unsigned scale = colors - 1;
    const bool small_map = colors <= 65536u;
    auto look = [&](unsigned v) -> unsigned {
        // render_tiles.rs:339-346: (value * (color_count - 1) + u16::MAX / 2) / u16::MAX  (32 bits hold it when colors <= 2^16)
        const unsigned ci = colors <= 1 ? 0u : (small_map ? (v * scale + 32767u) / 65535u
                                                          : static_cast<unsigned>((static_cast<unsigned long long>(v) * scale + 32767ull) / 65535ull));
        return __ldg(cm + ci);
    };
    for (unsigned x = 4 * threadIdx.x; x < width; x += 4 * kIdThreads) {
        const unsigned npx = min(4u, width - x);
#pragma unroll
        for (int r = 0; r < kIdRows; r++) {
            if (static_cast<unsigned>(r) >= rows) break;
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

}  // namespace

cudaError_t launch_spectrogram_tiles(const TileDesc *d_descs, int n, unsigned max_w, unsigned max_h, unsigned max_tmp_h,
                                     const uchar4 *d_colormap, unsigned colors, cudaStream_t st) {
    if (n <= 0 || !max_w || !max_h) return cudaSuccess;
    const unsigned gx = (max_w + kTileThreads - 1) / kTileThreads;
    for (int c0 = 0; c0 < n; c0 += 65535) {
        const unsigned nc = static_cast<unsigned>(n - c0 < 65535 ? n - c0 : 65535);
        // (each kernel returns at once for the descriptors of the other kind)
        tile_identity_kernel<<<dim3((max_h + kIdRows - 1) / kIdRows, nc), kIdThreads, 0, st>>>(d_descs + c0, d_colormap, colors);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
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
