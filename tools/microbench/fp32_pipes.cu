// fp32_pipes.cu -- B200 issue-rate microbenchmarks that decide the FFT kernel's instruction mix:
// FFMA / FADD / FMUL (3-register forms), the packed FFMA2 / FADD2 / FMUL2 forms of sm_100, SHFL,
// LDS.64 / LDS.128, MUFU.LG2.  Prints warp-instructions and lane-ops per clock per SM, measured
// with clock64() inside the kernel (1024 threads per SM, every SM busy).
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#define ITER 4096

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float *out, long long *cyc, float seed) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(seed + i, 1.f, 2.f, 3.f);
    __syncthreads();
    float a[8];
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; p[i] = make_float2(a[i], a[i] + 1.f); }
    const float b = seed * 0.5f + 1.0f, c = seed * 0.25f;
    const float2 b2 = make_float2(b, b + 0.5f), c2 = make_float2(c, c + 0.1f);
    int idx = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] + b;
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b;
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], b2, c2);
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __fadd2_rn(p[i], b2);
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __fmul2_rn(p[i], b2);
        } else if (MODE == 6) {  // shuffles
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
        } else if (MODE == 7) {  // LDS.64
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 v = *reinterpret_cast<const float2 *>(&sm[(idx + 32 * i) & 2047]);
                a[i] += v.x; idx += (int)v.y;
            }
        } else if (MODE == 8) {  // LDS.128
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 v = sm[(idx + 32 * i) & 2047];
                a[i] += v.x; idx += (int)v.y;
            }
        } else if (MODE == 9) {  // half FFMA, half FADD (butterfly-like mix)
#pragma unroll
            for (int i = 0; i < 8; i += 2) { a[i] = fmaf(a[i], b, c); a[i + 1] = a[i + 1] + b; }
#pragma unroll
            for (int i = 0; i < 8; i += 2) { a[i] = fmaf(a[i], b, c); a[i + 1] = a[i + 1] + b; }
        } else if (MODE == 10) {  // FFMA2 + FADD2 mix
#pragma unroll
            for (int i = 0; i < 8; i += 2) { p[i] = __ffma2_rn(p[i], b2, c2); p[i + 1] = __fadd2_rn(p[i + 1], b2); }
        } else if (MODE == 11) {  // MUFU.LG2
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __log2f(a[i]);
        } else if (MODE == 12) {  // FFMA + independent LDS.64 stream (does LDS steal FP32 issue slots?)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
            const float2 v = *reinterpret_cast<const float2 *>(&sm[(idx) & 2047]);
            p[0].x += v.x; idx += 32;
        } else if (MODE == 13) {  // FFMA2 + LDS.64 at the same ratio
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = __ffma2_rn(p[i], b2, c2);
            const float2 v = *reinterpret_cast<const float2 *>(&sm[(idx) & 2047]);
            p[7].x += v.x; idx += 32;
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + idx;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double warp_instr_per_iter, double lane_ops_per_instr, int sms, float *d_out, long long *d_cyc) {
    k<MODE><<<sms, 1024>>>(d_out, d_cyc, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, 1024>>>(d_out, d_cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= sms;
    const double winstr = 32.0 * ITER * warp_instr_per_iter;  // warp-instructions per SM (32 warps)
    printf("%-28s %8.0f cyc  %6.2f warp-instr/clk/SM  %7.1f lane-ops/clk/SM  (%.3f ms, err=%s)\n", name, avg,
           winstr / avg, winstr * 32.0 * lane_ops_per_instr / avg, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock %d MHz\n", p.name, sms, p.clockRate / 1000);
    float *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(float) * sms * 1024); cudaMalloc(&d_cyc, sizeof(long long) * sms);
    run<0>("FFMA", 8, 1, sms, d_out, d_cyc);
    run<1>("FADD", 8, 1, sms, d_out, d_cyc);
    run<2>("FMUL", 8, 1, sms, d_out, d_cyc);
    run<3>("FFMA2 (packed)", 8, 2, sms, d_out, d_cyc);
    run<4>("FADD2 (packed)", 8, 2, sms, d_out, d_cyc);
    run<5>("FMUL2 (packed)", 8, 2, sms, d_out, d_cyc);
    run<9>("FFMA+FADD 1:1", 8, 1, sms, d_out, d_cyc);
    run<10>("FFMA2+FADD2 1:1", 8, 2, sms, d_out, d_cyc);
    run<6>("SHFL.BFLY", 8, 1, sms, d_out, d_cyc);
    run<7>("LDS.64 (+FADD,+cvt)", 8, 1, sms, d_out, d_cyc);
    run<8>("LDS.128 (+FADD,+cvt)", 8, 1, sms, d_out, d_cyc);
    run<11>("MUFU.LG2", 8, 1, sms, d_out, d_cyc);
    run<12>("8 FFMA + 1 LDS.64", 8, 1, sms, d_out, d_cyc);
    run<13>("4 FFMA2 + 1 LDS.64", 4, 2, sms, d_out, d_cyc);
    return 0;
}
