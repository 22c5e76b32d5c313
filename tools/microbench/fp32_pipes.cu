// FP32 pipe microbenchmark for B200 (sm_100a): throughput and dependent-issue latency of the scalar and packed
// (f32x2) FP32 instructions the EQ and FFT kernels are built from, plus SHFL latency.  Not part of the product
// path: it records the FP32 issue peak used as the roofline denominator in DESIGN.md.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum Op { FADD, FMUL, FFMA, FADD2, FMUL2, FFMA2, MIXEQ, MIXEQ2 };

template <int OP> __device__ __forceinline__ void step(float (&a)[8], unsigned long long (&p)[8], float c, unsigned long long cc) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(c));
        if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc));
        if (OP == MIXEQ) {  // alternating mul / add like the DF2T cascade, scalar
            if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
            else asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        }
        if (OP == MIXEQ2) {
            if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
            else asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        }
    }
}

template <int OP> __global__ void tput(float* out, int iters, float c) {
    float a[8]; unsigned long long p[8];
    unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) step<OP>(a, p, c, cc);
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)(p[i] & 0xffffffffu)) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> __global__ void lat(float* out, long long* cyc, int iters, float c) {
    float a = threadIdx.x * 0.001f; unsigned long long p = __float_as_uint(a);
    unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64; ++u) {
            if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c));
            if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c));
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(c));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(cc));
            if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(cc));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(cc));
            if (OP == MIXEQ) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c)); }
            if (OP == MIXEQ2) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(cc)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(cc)); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = a + __uint_as_float((unsigned)p);
}

__global__ void shfl_lat(float* out, long long* cyc, int iters) {
    float a = threadIdx.x;
    int src = (threadIdx.x + 31) & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64; ++u) a = __shfl_sync(0xffffffffu, a, src);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = a;
}

template <int OP> int run(const char* name, int lanes_per_instr, int flops_per_lane, int sm_count, float mhz) {
    float* out; long long* cyc;
    const int blocks = sm_count * 8, threads = 256, iters = 4096;
    CK(cudaMalloc(&out, sizeof(float) * blocks * threads)); CK(cudaMalloc(&cyc, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    tput<OP><<<blocks, threads>>>(out, 64, 1.0001f);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); tput<OP><<<blocks, threads>>>(out, iters, 1.0001f); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double instr = (double)blocks * threads / 32 * iters * 64.0;          // warp instructions
    double wips = instr / (best * 1e-3);
    double per_sm_clk = wips / sm_count / (mhz * 1e6);
    double tflops = wips * 32 * lanes_per_instr * flops_per_lane / 1e12;
    lat<OP><<<1, 32>>>(out, cyc, 64, 1.0001f); CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    double per_op = (double)c / (64.0 * 64 * ((OP == MIXEQ || OP == MIXEQ2) ? 2 : 1));
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"warp_instr_per_s\": %.4e, \"warp_instr_per_sm_per_clk_at_%dMHz\": %.3f, \"tflops\": %.2f, \"dep_latency_cyc\": %.2f}\n",
           name, best, wips, (int)mhz, per_sm_clk, tflops, per_op);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    float mhz = clk_khz / 1000.f;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f}\n", prop.name, prop.multiProcessorCount, mhz);
    int n = prop.multiProcessorCount;
    if (run<FADD>("add.f32", 1, 1, n, mhz)) return 1;
    if (run<FMUL>("mul.f32", 1, 1, n, mhz)) return 1;
    if (run<FFMA>("fma.f32", 1, 2, n, mhz)) return 1;
    if (run<FADD2>("add.f32x2", 2, 1, n, mhz)) return 1;
    if (run<FMUL2>("mul.f32x2", 2, 1, n, mhz)) return 1;
    if (run<FFMA2>("fma.f32x2", 2, 2, n, mhz)) return 1;
    if (run<MIXEQ>("mul+add.f32 alternating", 1, 1, n, mhz)) return 1;
    if (run<MIXEQ2>("mul+add.f32x2 alternating", 2, 1, n, mhz)) return 1;
    float* out; long long* cyc; CK(cudaMalloc(&out, 128)); CK(cudaMalloc(&cyc, 8));
    shfl_lat<<<1, 32>>>(out, cyc, 64); CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"op\": \"shfl.sync.idx\", \"dep_latency_cyc\": %.2f}\n", (double)c / (64.0 * 64));
    return 0;
}
