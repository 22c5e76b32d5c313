// Single-warp issue rates on B200 (sm_100a): how fast ONE warp per scheduler partition issues independent scalar FP32
// instructions against packed f32x2 ones, and what a DF2T biquad step costs in the two forms the EQ warps can take
// (scalar: two bands per lane, one channel; packed: one band per lane, left and right in one f32x2).  Not part of the
// product path: it backs the choice of the EQ warps' arithmetic in DESIGN.md.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ void unpk(u64 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// NI independent chains of one op per warp; warps_per_block warps (4 = one per partition, 8 = two ...)
template <int OP, int NI> __global__ void indep(float* out, long long* cyc, int iters, float c) {
    float a[NI]; u64 p[NI];
    const u64 cc = pk(c, c);
#pragma unroll
    for (int i = 0; i < NI; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 1.f); }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if (OP == 0) { if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c)); else asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c)); }
                if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc));
                if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(c));
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    float s = 0;
#pragma unroll
    for (int i = 0; i < NI; ++i) { float x, y; unpk(p[i], x, y); s += a[i] + x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// scalar DF2T step (reference operation order, never contracted)
__device__ __forceinline__ float df2t(float x, float& s1, float& s2, float b0, float b1, float b2, float a1, float a2) {
    const float out = __fadd_rn(s1, __fmul_rn(b0, x));
    s1 = __fsub_rn(__fadd_rn(s2, __fmul_rn(b1, x)), __fmul_rn(a1, out));
    s2 = __fsub_rn(__fmul_rn(b2, x), __fmul_rn(a2, out));
    return out;
}
// packed DF2T step: every product is fma(a, b, -0) and every sum fma(a, 1, c) with run-time 1 and -0, so each op is
// rounded once, exactly like the scalar mul/add/sub
__device__ __forceinline__ u64 df2t2(u64 x, u64& s1, u64& s2, u64 b0, u64 b1, u64 b2, u64 na1, u64 na2, u64 one, u64 nz) {
    const u64 out = fma2(fma2(b0, x, nz), one, s1);
    s1 = fma2(fma2(fma2(b1, x, nz), one, s2), one, fma2(na1, out, nz));
    s2 = fma2(fma2(b2, x, nz), one, fma2(na2, out, nz));
    return out;
}

// the systolic loop, scalar, two bands per lane (the product's present form, simplified: no loads or stores)
__global__ void eq_scalar(float* out, long long* cyc, int steps, const float* cf) {
    float ab0 = cf[0], ab1 = cf[1], ab2 = cf[2], aa1 = cf[3], aa2 = cf[4], bb0 = cf[5], bb1 = cf[6], bb2 = cf[7], ba1 = cf[8], ba2 = cf[9];
    float as1 = 0, as2 = 0, bs1 = 0, bs2 = 0, ya = 0, xs[4] = {0.1f, 0.2f, 0.3f, 0.4f};
    const int src = (threadIdx.x + 26) & 31;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < steps; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float y = df2t(ya, bs1, bs2, bb0, bb1, bb2, ba1, ba2);
            ya = df2t(xs[u], as1, as2, ab0, ab1, ab2, aa1, aa2);
            xs[(u + 3) % 4] = __shfl_sync(0xffffffffu, y, src);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = as1 + as2 + bs1 + bs2 + ya + xs[0];
}
// scalar, one band per lane (config 5's present form)
__global__ void eq_scalar1(float* out, long long* cyc, int steps, const float* cf) {
    float ab0 = cf[0], ab1 = cf[1], ab2 = cf[2], aa1 = cf[3], aa2 = cf[4];
    float as1 = 0, as2 = 0, xs[8] = {0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f};
    const int src = (threadIdx.x + 29) & 31;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < steps; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float y = df2t(xs[u], as1, as2, ab0, ab1, ab2, aa1, aa2);
            xs[(u + 7) % 8] = __shfl_sync(0xffffffffu, y, src);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = as1 + as2 + xs[0];
}
// packed, one band per lane, left and right in one f32x2; DL = lane skew in steps
template <int DL> __global__ void eq_packed(float* out, long long* cyc, int steps, const float* cf, float one_f, float nz_f) {
    const u64 one = pk(one_f, one_f), nz = pk(nz_f, nz_f);
    const u64 b0 = pk(cf[0], cf[0]), b1 = pk(cf[1], cf[1]), b2 = pk(cf[2], cf[2]), na1 = pk(-cf[3], -cf[3]), na2 = pk(-cf[4], -cf[4]);
    u64 s1 = 0, s2 = 0, xs[DL];
#pragma unroll
    for (int u = 0; u < DL; ++u) xs[u] = pk(0.1f * u, 0.2f * u);
    const int src = (threadIdx.x + 29) & 31;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < steps; i += DL) {
#pragma unroll
        for (int u = 0; u < DL; ++u) {
            const u64 y = df2t2(xs[u], s1, s2, b0, b1, b2, na1, na2, one, nz);
            xs[(u + DL - 1) % DL] = __shfl_sync(0xffffffffu, y, src);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    float x, y; unpk(s1 ^ s2 ^ xs[0], x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}

int main() {
    float* out; long long* cyc; float* cf;
    CK(cudaMalloc(&out, 1 << 20)); CK(cudaMalloc(&cyc, 4096)); CK(cudaMalloc(&cf, 64));
    const float hcf[10] = {1.01f, -1.9f, 0.95f, -1.9f, 0.94f, 0.99f, -1.8f, 0.9f, -1.8f, 0.89f};
    CK(cudaMemcpy(cf, hcf, sizeof hcf, cudaMemcpyHostToDevice));
    long long h[64];
    const int iters = 4096;
    auto report = [&](const char* name, int warps, double per, const char* unit) {
        cudaDeviceSynchronize(); cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("{\"test\": \"%s\", \"warps_per_sm\": %d, \"cycles_per_%s\": %.3f}\n", name, warps, unit, (double)mx / per);
    };
    for (int warps : {4, 8, 16}) {
        indep<0, 8><<<1, 32 * warps>>>(out, cyc, iters, 1.0001f); report("scalar mul/add, 8 independent chains per warp", warps, iters * 64.0, "instr");
        indep<2, 8><<<1, 32 * warps>>>(out, cyc, iters, 1.0001f); report("scalar fma, 8 independent chains per warp", warps, iters * 64.0, "instr");
        indep<1, 8><<<1, 32 * warps>>>(out, cyc, iters, 1.0001f); report("fma.f32x2, 8 independent chains per warp", warps, iters * 64.0, "instr");
    }
    for (int warps : {1, 4, 8}) {
        const int steps = 1 << 16;
        eq_scalar<<<1, 32 * warps>>>(out, cyc, steps, cf); report("DF2T scalar, two bands per lane, skew 4", warps, steps, "step");
        eq_scalar1<<<1, 32 * warps>>>(out, cyc, steps, cf); report("DF2T scalar, one band per lane, skew 8", warps, steps, "step");
        eq_packed<8><<<1, 32 * warps>>>(out, cyc, steps, cf, 1.0f, -0.0f); report("DF2T packed L/R, one band per lane, skew 8", warps, steps, "step");
        eq_packed<4><<<1, 32 * warps>>>(out, cyc, steps, cf, 1.0f, -0.0f); report("DF2T packed L/R, one band per lane, skew 4", warps, steps, "step");
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
