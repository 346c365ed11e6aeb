// Micro-benchmark of the issue / pipe rates that bound the render kernels (run on the B200 box):
//   FFMA (3 distinct registers), FFMA2 (packed fp32x2), FMNMX (alu pipe), MUFU.RCP / MUFU.EX2, and mixes.
// Each kernel runs `iters` unrolled rounds of N independent chains per thread; every SM gets `ctas`
// CTAs of 256 threads so that all four schedulers have 2*ctas warps.  Prints warp-instructions per
// cycle per SM sub-partition for each mix (1.0 = the issue limit).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rcpa(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2a(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int N = 8;   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[N]; u64 y[N];
    for (int i = 0; i < N; i++) { x[i] = a + i + threadIdx.x; y[i] = pk(x[i], b + i); }
    const u64 pa = pk(a, b), pb = pk(b, a);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (MODE == 0) x[i] = fmaf(x[i], a, b);                                 // FFMA
                if (MODE == 1) y[i] = fma2(y[i], pa, pb);                                // FFMA2
                if (MODE == 2) x[i] = fminf(x[i], x[(i + 1) % N]);                       // FMNMX (alu pipe)
                if (MODE == 3) x[i] = ex2a(x[i]);                                        // MUFU.EX2
                if (MODE == 4) { x[i] = fmaf(x[i], a, b); if ((i & 3) == 0) x[i] = ex2a(x[i]); }   // 4 FFMA : 1 MUFU
                if (MODE == 5) { x[i] = fmaf(x[i], a, b); x[i] = fminf(x[i], x[(i + 1) % N]); }   // FFMA + FMNMX alternating
                if (MODE == 6) { y[i] = fma2(y[i], pa, pb); x[i] = fminf(x[i], x[(i + 1) % N]); }  // FFMA2 + FMNMX alternating
            }
        }
    }
    float s = 0;
    for (int i = 0; i < N; i++) { float p, q; upk(y[i], p, q); s += x[i] + p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128 broadcast patterns: all 32 lanes one address / two half-warp addresses / four addresses
template <int GROUPS>
__global__ void __launch_bounds__(256) lds(float* out, int iters) {
    __shared__ float4 s[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) s[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = (lane / (32 / GROUPS)) * 37 + (threadIdx.x >> 5);
    float acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const float4 v = s[(idx + u * 5) & 1023];
            acc += (v.x + v.y) + (v.z + v.w);
        }
        idx += (int)acc & 1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, ctas = 4, iters = 2000;
    float* out; cudaMalloc(&out, (size_t)sms * ctas * 256 * 4);
    printf("device %s, %d SMs, max clock %.0f MHz\n", p.name, sms, clk_khz / 1000.0);
    const char* names[] = {"FFMA", "FFMA2", "FMNMX", "MUFU.EX2", "4 FFMA : 1 MUFU.EX2", "FFMA+FMNMX", "FFMA2+FMNMX"};
    const double per_iter[] = {8.0 * N, 8.0 * N, 8.0 * N, 8.0 * N, 8.0 * N * 1.25, 16.0 * N, 16.0 * N};
    for (int m = 0; m < 7; m++) {
        float ms = 0;
        auto run = [&] {
            switch (m) {
                case 0: k<0><<<sms * ctas, 256>>>(out, iters, 1.0001f, 0.5f); break;
                case 1: k<1><<<sms * ctas, 256>>>(out, iters, 1.0001f, 0.5f); break;
                case 2: k<2><<<sms * ctas, 256>>>(out, iters, 1.0001f, 0.5f); break;
                case 3: k<3><<<sms * ctas, 256>>>(out, iters, 1.0001f, 0.5f); break;
                case 4: k<4><<<sms * ctas, 256>>>(out, iters, 1.0001f, 0.5f); break;
                case 5: k<5><<<sms * ctas, 256>>>(out, iters, 1.0001f, 0.5f); break;
                case 6: k<6><<<sms * ctas, 256>>>(out, iters, 1.0001f, 0.5f); break;
            }
        };
        ms = time_ms(run);
        const double warp_instr = (double)sms * ctas * 8 * iters * per_iter[m];
        // cycles at the max clock (the box may run below it; compare ratios between rows)
        const double cycles = ms * 1e-3 * clk_khz * 1e3;
        printf("%-22s %8.3f ms  %.3f warp-instr / cycle / SMSP (at max clock)\n", names[m], ms, warp_instr / (cycles * sms * 4));
    }
    for (int g = 1; g <= 4; g *= 2) {
        float ms = time_ms([&] {
            if (g == 1) lds<1><<<sms * ctas, 256>>>(out, iters);
            if (g == 2) lds<2><<<sms * ctas, 256>>>(out, iters);
            if (g == 4) lds<4><<<sms * ctas, 256>>>(out, iters);
        });
        const double n = (double)sms * ctas * 8 * iters * 16;
        const double cycles = ms * 1e-3 * clk_khz * 1e3;
        printf("LDS.128 %d address(es)   %8.3f ms  %.3f LDS / cycle / SM\n", g, ms, n / (cycles * sms));
    }
    return 0;
}
