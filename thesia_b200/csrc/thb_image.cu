// thb_image.cu -- K3 tail (global {max,-min} reduce + clamp rules) and K4 (dB -> u16, transposed).
//
// K4 = convert_spectrogram_to_img (visualize/drawing.rs:4-33): out[i][j] = q(spec[j][i0 + i]),
// q(v) = clamp(round((v - min_dB) / (max_dB - min_dB) * (65535 - min_value) + min_value), 0, 65535)
// with min_value = max(1, round(65535 / colormap_length)); rows >= B are 0; (-inf, -inf) range
// gives an all-zero image.  Multiply and add are NOT fused (Rust never contracts), so a pixel is
// bit-identical to the reference's whenever the dB input is.
// HBM-bound: reads 4 B and writes 2 B per pixel; a 32(bins) x 64(frames) shared-memory tile turns
// the reference's strided gather into coalesced 128 B reads and 128 B writes.
#include "thb_device.cuh"
#include "thb_kernels.cuh"

namespace thb {
namespace {

__global__ void minmax_init_kernel(float *slots, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * n) slots[i] = -CUDART_INF_F;
}

__global__ void minmax_init_tracks_kernel(const TrackDesc *__restrict__ tracks, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        tracks[i].minmax[0] = -CUDART_INF_F;
        tracks[i].minmax[1] = -CUDART_INF_F;
    }
}

// grid-stride {max, -min} of a flat array: warp shuffle -> one atomic pair per warp
__global__ void __launch_bounds__(256) minmax_array_kernel(const float *__restrict__ x, unsigned long long n,
                                                           float *slot) {
    float a = -CUDART_INF_F, b = -CUDART_INF_F;
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __ldg(x + i);
        a = fmaxf(a, v);
        b = fmaxf(b, -v);
    }
    a = warp_max(a);
    b = warp_max(b);
    if ((threadIdx.x & 31) == 0) {
        atomic_max_float(&slot[0], a);
        atomic_max_float(&slot[1], b);
    }
}

// one warp: max over all slots of {max, -min}  (the rayon reduce of mod.rs:169-178)
__global__ void minmax_reduce_kernel(const float *__restrict__ slots, int n_slots, float *send) {
    pdl_wait();
    float a = -CUDART_INF_F, b = -CUDART_INF_F;
    for (int i = threadIdx.x; i < n_slots; i += 32) {
        a = fmaxf(a, slots[2 * i]);
        b = fmaxf(b, slots[2 * i + 1]);
    }
    a = warp_max(a);
    b = warp_max(b);
    if (threadIdx.x == 0) {
        send[0] = a;
        send[1] = b;
    }
}

// The global dB range in ONE kernel when the ranks of a box can see each other's memory (thb_comm_init maps every
// peer's exchange slots through CUDA IPC over NVLink): local max over the slots, each lane l < n_ranks stores
// {max, -min} + the step's sequence number into ITS slot of peer l (remote store, payload before flag), polls peer l's
// slot in the local buffer until the same sequence number arrives, max over ranks, clamp rules (mod.rs:179-180).
// Replaces minmax_reduce + ncclAllReduce(max, 2 floats) + minmax_finalize: three launches and NCCL's ~40 us for 8 bytes
// become one launch and an NVLink round trip.  Slots are double buffered by the parity of the sequence number: a rank
// can start step s + 1 (other parity) while a slow rank still reads step s, and nobody can reach step s + 2 before
// every rank has finished reading step s.  `fail` is set (and the range left NaN) if a peer does not show up in ~5 s.
__global__ void minmax_exchange_kernel(const float *__restrict__ slots, int n_slots, float4 *const *peers, float4 *mine, int n_ranks,
                                       int rank, unsigned seq, float dB_range, float *range, float *send, unsigned *fail) {
    pdl_wait();
    pdl_launch_dependents();   // the quantiser's CTAs may take their places while this warp waits for the peers
    const int lane = threadIdx.x;
    float a = -CUDART_INF_F, b = -CUDART_INF_F;
    for (int i = lane; i < n_slots; i += 32) {
        a = fmaxf(a, slots[2 * i]);
        b = fmaxf(b, slots[2 * i + 1]);
    }
    a = warp_max(a);
    b = warp_max(b);
    const int base = (seq & 1u) * kExchangeMaxRanks;
    if (lane < n_ranks) {
        volatile float *dst = reinterpret_cast<volatile float *>(peers[lane] + base + rank);
        dst[0] = a;
        dst[1] = b;
        __threadfence_system();
        reinterpret_cast<volatile unsigned *>(dst)[2] = seq;
    }
    float ra = -CUDART_INF_F, rb = -CUDART_INF_F;
    bool ok = true;
    if (lane < n_ranks) {
        volatile float *src = reinterpret_cast<volatile float *>(mine + base + lane);
        unsigned spins = 0;
        while (reinterpret_cast<volatile unsigned *>(src)[2] != seq) {
            __nanosleep(100);
            if (++spins > 50000000u) {
                ok = false;
                break;
            }
        }
        __threadfence_system();
        ra = src[0];
        rb = src[1];
    }
    ok = __all_sync(0xffffffffu, ok);
    ra = warp_max(ra);
    rb = warp_max(rb);
    if (lane == 0) {
        if (!ok) {
            *fail = 1u;
            range[0] = range[1] = CUDART_NAN_F;
            return;
        }
        send[0] = ra;
        send[1] = rb;
        const float mx = fminf(ra, 0.0f);
        range[0] = fmaxf(-rb, __fsub_rn(mx, dB_range));
        range[1] = mx;
    }
}

// max <- min(max, 0); min <- max(min, max - dB_range)   (mod.rs:179-180)
__global__ void minmax_finalize_kernel(const float *__restrict__ send, float dB_range, float *range) {
    pdl_wait();
    const float mx = fminf(send[0], 0.0f);
    const float mn = fmaxf(-send[1], __fsub_rn(mx, dB_range));
    range[0] = mn;
    range[1] = mx;
}

constexpr int kTileT = 64;  // frames per tile (image columns)
constexpr int kTileB = 32;  // bins per tile   (image rows)

__global__ void __launch_bounds__(256) spec_to_img_kernel(const ImgDesc *__restrict__ descs,
                                                          const float *__restrict__ range,
                                                          float min_value_f, float u16_span) {
    __shared__ uint16_t tile[kTileB][kTileT + 2];
    pdl_wait();   // launched under the range kernel: d_range (and, in stream order, the spectrograms) are ready from here on
    const ImgDesc d = descs[blockIdx.z];
    const long long t0 = static_cast<long long>(blockIdx.x) * kTileT;
    const int r0 = blockIdx.y * kTileB;  // image row (relative to i0)
    if (t0 >= d.T || r0 >= d.H) return;
    const float dB_min = range[0], dB_max = range[1];
    const bool all_zero = (dB_min == dB_max) && (dB_max == -CUDART_INF_F);
    const float span = __fsub_rn(dB_max, dB_min);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps
    // read: each warp takes frames warp, warp+8, ...; lanes run along the bins axis (contiguous)
#pragma unroll
    for (int i = 0; i < kTileT / 8; i++) {
        const int tt = warp + 8 * i;
        const long long t = t0 + tt;
        const int bin = d.i0 + r0 + lane;
        uint16_t px = 0;
        if (t < d.T && (r0 + lane) < d.H && bin < d.B && !all_zero) {
            const float v = __ldg(&d.spec[t * d.B + bin]);
            const float zero_to_one = __fdiv_rn(__fsub_rn(v, dB_min), span);
            const float scaled = __fadd_rn(__fmul_rn(zero_to_one, u16_span), min_value_f);
            const float r = roundf(scaled);  // f32::round: half away from zero
            // clamp(0, 65535) then `as u16`; NaN -> 0
            px = (r != r) ? 0 : static_cast<uint16_t>(fminf(fmaxf(r, 0.0f), 65535.0f));
        }
        tile[lane][tt] = px;
    }
    __syncthreads();
    // write: each warp takes image rows warp, warp+8, ...; lanes run along time, 2 pixels each
#pragma unroll
    for (int i = 0; i < kTileB / 8; i++) {
        const int rr = warp + 8 * i;
        if (r0 + rr >= d.H) continue;
        uint16_t *orow = d.img + static_cast<long long>(r0 + rr) * d.pitch;
        const long long t = t0 + 2 * lane;
        if (t + 1 < d.T && ((reinterpret_cast<uintptr_t>(orow + t) & 3) == 0)) {
            const uint32_t two = static_cast<uint32_t>(tile[rr][2 * lane]) |
                                 (static_cast<uint32_t>(tile[rr][2 * lane + 1]) << 16);
            *reinterpret_cast<uint32_t *>(orow + t) = two;
        } else {
            if (t < d.T) orow[t] = tile[rr][2 * lane];
            if (t + 1 < d.T) orow[t + 1] = tile[rr][2 * lane + 1];
        }
    }
}

// The quantiser of drawing.rs:22-31, operation for operation (no contraction).
__device__ __forceinline__ uint32_t quantise_dB(float v, float dB_min, float span, float u16_span, float min_value_f) {
    const float zero_to_one = __fdiv_rn(__fsub_rn(v, dB_min), span);
    const float scaled = __fadd_rn(__fmul_rn(zero_to_one, u16_span), min_value_f);
    const float r = roundf(scaled);  // f32::round: half away from zero
    return (r != r) ? 0u : static_cast<uint32_t>(fminf(fmaxf(r, 0.0f), 65535.0f));  // clamp, `as u16`; NaN -> 0
}

// The same quantiser with the division by the (launch-uniform) span replaced by Markstein's sequence on its correctly
// rounded reciprocal y = RN(1 / span):
//     q0 = RN(a y);  r0 = a - span q0 (exact, one FMA);  q1 = RN(q0 + r0 y)       -> q1 is a faithful quotient
//     r1 = a - span q1 (exact);                           q  = RN(q1 + r1 y)       -> q == RN(a / span)
// (Markstein 1990; Muller et al., Handbook of Floating-Point Arithmetic, "division with an FMA": with y the correctly
// rounded reciprocal and q1 faithful, the last step yields the correctly rounded quotient when nothing over- or
// underflows.)  One FMUL + four FFMA instead of __fdiv_rn's reciprocal refinement + range check + slow path, and the
// pixel is still the reference's bit for bit.  The caller guarantees 2^-60 < span < 2^60 (else the plain kernel runs);
// |a| >= 1e30 (incl. +-inf, NaN) takes a y alone: far outside the image range, it clamps -- or stays NaN -- exactly like
// the true quotient.  round(): trunc + compare, exact for every x that survives the clamp to [0, 65535]
// (x < 0 ends at <= 0, NaN at 0 through fmaxf, like f32::round followed by the reference's clamp and `as u16`).
__device__ __forceinline__ uint32_t quantise_dB_rcp(float v, float dB_min, float span, float y, float u16_span, float min_value_f) {
    const float a = __fsub_rn(v, dB_min);
    const float q0 = __fmul_rn(a, y);
    const float r0 = __fmaf_rn(-span, q0, a);
    const float q1 = __fmaf_rn(r0, y, q0);
    const float r1 = __fmaf_rn(-span, q1, a);
    const float q2 = __fmaf_rn(r1, y, q1);
    const float zero_to_one = fabsf(a) < 1e30f ? q2 : q0;
    const float scaled = __fadd_rn(__fmul_rn(zero_to_one, u16_span), min_value_f);
    const float t = truncf(scaled);
    const float r = (__fsub_rn(scaled, t) >= 0.5f) ? __fadd_rn(t, 1.0f) : t;
    return min(__float2uint_rz(r), 65535u);  // the conversion saturates: negative -> 0, NaN -> 0, +inf -> 2^32 - 1
}

// 128 (bins) x 128 (frames) tiles: a thread quantises 4 consecutive bins of 2 consecutive frames at a time (two
// 16-byte loads when VEC), packs the two frames of a bin into one 32-bit word and parks it in a shared tile whose
// columns are XOR-swizzled by the lane, so both the transposing stores and the row reads are conflict-free; image
// rows leave as 128-byte warp stores.
constexpr int kBigT = 128, kBigB = 128;

// RCP: quantise_dB_rcp (the launcher checks the span)
template <bool VEC, bool RCP>
__global__ void __launch_bounds__(256) spec_to_img_tile_kernel(const ImgDesc *__restrict__ descs,
                                                                const float *__restrict__ range, float min_value_f,
                                                                float u16_span) {
    __shared__ uint32_t tile[kBigB][kBigT / 2];
    pdl_wait();   // launched under the range kernel: d_range (and, in stream order, the spectrograms) are ready from here on
    const ImgDesc d = descs[blockIdx.z];
    const long long t0 = static_cast<long long>(blockIdx.x) * kBigT;
    const int r0 = blockIdx.y * kBigB;
    if (t0 >= d.T || r0 >= d.H) return;
    const float dB_min = range[0], dB_max = range[1];
    const bool all_zero = (dB_min == dB_max) && (dB_max == -CUDART_INF_F);
    const float span = __fsub_rn(dB_max, dB_min);
    // RCP needs an ordinary span; anything else (a zero, huge or non-finite range) takes the plain division
    const bool use_rcp = RCP && span > 8.6736174e-19f && span < 1.1529215e+18f;
    const float y = use_rcp ? __frcp_rn(span) : 0.0f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = r0 + 4 * lane;      // first image row of this lane's 4 bins
    const int bin = d.i0 + row;
    // A tile that lies wholly inside the spectrogram and the image (all but the last row / column of tiles) needs none
    // of the per-pixel bounds predicates: two 16-byte loads, eight quantisations, four packed words per step.  The
    // kernel is issue-bound (ncu: 72 % of the issue slots against 44 % of DRAM cycles, profiles/r02_img_kernel_ncu.txt),
    // so the instructions saved here are time saved.
    if (VEC && RCP && use_rcp && !all_zero && t0 + kBigT <= d.T && r0 + kBigB <= d.H && d.i0 + r0 + kBigB <= d.B) {
        const float *src = d.spec + (t0 + 2 * warp) * d.B + bin;
        const long long step = 16ll * d.B;  // eight frame pairs further
        uint32_t *trow = &tile[4 * lane][0];
#pragma unroll 4
        for (int i = 0; i < kBigT / 16; i++) {
            const int fp = warp + 8 * i;
            const float4 qa = __ldg(reinterpret_cast<const float4 *>(src + i * step));
            const float4 qb = __ldg(reinterpret_cast<const float4 *>(src + i * step + d.B));
            uint32_t *tcol = trow + (fp ^ lane);
            tcol[0 * (kBigT / 2)] = __byte_perm(quantise_dB_rcp(qa.x, dB_min, span, y, u16_span, min_value_f),
                                                quantise_dB_rcp(qb.x, dB_min, span, y, u16_span, min_value_f), 0x5410);
            tcol[1 * (kBigT / 2)] = __byte_perm(quantise_dB_rcp(qa.y, dB_min, span, y, u16_span, min_value_f),
                                                quantise_dB_rcp(qb.y, dB_min, span, y, u16_span, min_value_f), 0x5410);
            tcol[2 * (kBigT / 2)] = __byte_perm(quantise_dB_rcp(qa.z, dB_min, span, y, u16_span, min_value_f),
                                                quantise_dB_rcp(qb.z, dB_min, span, y, u16_span, min_value_f), 0x5410);
            tcol[3 * (kBigT / 2)] = __byte_perm(quantise_dB_rcp(qa.w, dB_min, span, y, u16_span, min_value_f),
                                                quantise_dB_rcp(qb.w, dB_min, span, y, u16_span, min_value_f), 0x5410);
        }
    } else
#pragma unroll 2
    for (int i = 0; i < kBigT / 16; i++) {
        const int fp = warp + 8 * i;    // frame pair inside the tile
        const long long t = t0 + 2 * fp;
        float v[2][4];
        bool ok[2][4];
#pragma unroll
        for (int f = 0; f < 2; f++) {
            const bool t_ok = (t + f) < d.T && !all_zero;
            const float *src = d.spec + (t + f) * d.B + bin;
            if (VEC && t_ok && row + 3 < d.H && bin + 3 < d.B) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(src));
                v[f][0] = q.x; v[f][1] = q.y; v[f][2] = q.z; v[f][3] = q.w;
                ok[f][0] = ok[f][1] = ok[f][2] = ok[f][3] = true;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    ok[f][j] = t_ok && (row + j) < d.H && (bin + j) < d.B;
                    v[f][j] = ok[f][j] ? __ldg(src + j) : 0.0f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t a, b;
            if (use_rcp) {
                a = ok[0][j] ? quantise_dB_rcp(v[0][j], dB_min, span, y, u16_span, min_value_f) : 0u;
                b = ok[1][j] ? quantise_dB_rcp(v[1][j], dB_min, span, y, u16_span, min_value_f) : 0u;
            } else {
                a = ok[0][j] ? quantise_dB(v[0][j], dB_min, span, u16_span, min_value_f) : 0u;
                b = ok[1][j] ? quantise_dB(v[1][j], dB_min, span, u16_span, min_value_f) : 0u;
            }
            tile[4 * lane + j][fp ^ lane] = a | (b << 16);
        }
    }
    __syncthreads();
    if (r0 + kBigB <= d.H && t0 + kBigT <= d.pitch) {
        // full tile: no bounds to check (pitch is a multiple of 64 frames and rows are 4-byte aligned in this mode)
        uint16_t *obase = d.img + static_cast<long long>(r0 + warp) * d.pitch + t0 + 2 * lane;
        const long long ostep = 8ll * d.pitch;
#pragma unroll 8
        for (int i = 0; i < kBigB / 8; i++) {
            const int rr = warp + 8 * i;
            const int g = (rr >> 2) & 31;
            uint32_t *o = reinterpret_cast<uint32_t *>(obase + i * ostep);
            o[0] = tile[rr][lane ^ g];
            o[32] = tile[rr][(lane + 32) ^ g];
        }
        return;
    }
#pragma unroll 4
    for (int i = 0; i < kBigB / 8; i++) {
        const int rr = warp + 8 * i;
        if (r0 + rr >= d.H) break;
        const int g = (rr >> 2) & 31;
        uint16_t *orow = d.img + static_cast<long long>(r0 + rr) * d.pitch + t0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = lane + 32 * h;
            const uint32_t w = tile[rr][c ^ g];
            const long long t = t0 + 2 * c;
            if (t + 1 < d.pitch) {
                *reinterpret_cast<uint32_t *>(orow + 2 * c) = w;   // columns >= T inside the pitch get zeros
            } else if (t < d.T) {
                orow[2 * c] = static_cast<uint16_t>(w & 0xffffu);
            }
        }
    }
}

}  // namespace

cudaError_t launch_minmax_init(float *d_slots, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    minmax_init_kernel<<<(2 * n + 255) / 256, 256, 0, st>>>(d_slots, n);
    return cudaGetLastError();
}

cudaError_t launch_minmax_init_tracks(const TrackDesc *d_tracks, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    minmax_init_tracks_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_tracks, n);
    return cudaGetLastError();
}

cudaError_t launch_minmax_array(const float *d_x, unsigned long long n, float *d_slot, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long blocks = (n + 255) / 256;
    const unsigned long long cap = static_cast<unsigned long long>(sm_count) * 8;
    if (blocks > cap) blocks = cap;
    minmax_array_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(d_x, n, d_slot);
    return cudaGetLastError();
}

cudaError_t launch_minmax_reduce(const float *d_slots, int n_slots, float *d_send, cudaStream_t st) {
    return launch_pdl(minmax_reduce_kernel, dim3(1), dim3(32), 0, st, d_slots, n_slots, d_send);
}

cudaError_t launch_minmax_exchange(const float *d_slots, int n_slots, float4 *const *d_peers, float4 *d_mine, int n_ranks, int rank,
                                   unsigned seq, float dB_range, float *d_range, float *d_send, unsigned *d_fail, cudaStream_t st) {
    return launch_pdl(minmax_exchange_kernel, dim3(1), dim3(32), 0, st, d_slots, n_slots, d_peers, d_mine, n_ranks, rank, seq, dB_range, d_range,
                      d_send, d_fail);
}

cudaError_t launch_minmax_finalize(const float *d_send, float dB_range, float *d_range, cudaStream_t st) {
    return launch_pdl(minmax_finalize_kernel, dim3(1), dim3(1), 0, st, d_send, dB_range, d_range);
}

cudaError_t launch_spec_to_img(const ImgDesc *d_descs, int n, long long max_T, int max_H,
                               const float *d_range, uint32_t colormap_length, int tile_mode, cudaStream_t st) {
    if (n <= 0 || max_T <= 0 || max_H <= 0) return cudaSuccess;
    // min_value = max(1, round(65535 / colormap_length) as u16)   (drawing.rs:20-21)
    double r = colormap_length ? std::round(65535.0 / static_cast<double>(colormap_length)) : 65535.0;
    if (r > 65535.0) r = 65535.0;
    uint32_t min_value = static_cast<uint32_t>(r);
    if (min_value < 1) min_value = 1;
    const float u16_span = static_cast<float>(65535u - min_value);
    // tile_mode: 0 = small tiles (any layout), 1 = big tiles, 2 = big tiles with 16-byte loads, 3 = as 2 with the
    // reciprocal-based exact quotient
    const unsigned gx = static_cast<unsigned>(tile_mode ? (max_T + kBigT - 1) / kBigT : (max_T + kTileT - 1) / kTileT);
    const unsigned gy = static_cast<unsigned>(tile_mode ? (max_H + kBigB - 1) / kBigB : (max_H + kTileB - 1) / kTileB);
    for (int t0 = 0; t0 < n; t0 += 65535) {
        const int nt = n - t0 < 65535 ? n - t0 : 65535;
        dim3 grid(gx, gy, static_cast<unsigned>(nt));
        const float mv = static_cast<float>(min_value);
        cudaError_t e;
        if (tile_mode == 3) e = launch_pdl(spec_to_img_tile_kernel<true, true>, grid, dim3(256), 0, st, d_descs + t0, d_range, mv, u16_span);
        else if (tile_mode == 2) e = launch_pdl(spec_to_img_tile_kernel<true, false>, grid, dim3(256), 0, st, d_descs + t0, d_range, mv, u16_span);
        else if (tile_mode == 1) e = launch_pdl(spec_to_img_tile_kernel<false, false>, grid, dim3(256), 0, st, d_descs + t0, d_range, mv, u16_span);
        else e = launch_pdl(spec_to_img_kernel, grid, dim3(256), 0, st, d_descs + t0, d_range, mv, u16_span);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace thb
