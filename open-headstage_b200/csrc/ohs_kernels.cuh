// ohs_kernels.cuh — hand-written sm_100a kernels of the Open Headstage DSP hot path.
//
// One fused kernel renders, for a group of G stereo streams per CTA and K engine blocks per launch,
//     10-band DF2T biquad cascade (bit-exact)  ->  uniformly partitioned overlap-save FFT convolution against the four
//     HRIR paths  ->  ear sums  ->  output gain
// i.e. the work of reference src/lib.rs:1179-1207 (Plugin::process) over src/dsp/parametric_eq.rs:166-179 and
// src/dsp/convolution.rs:184-289, for many streams at once.
//
// Design (see DESIGN.md for the derivations and the measured pipe numbers behind them):
//   * warp 0 of the CTA is the EQ warp: band-systolic — lane (g, j) runs band j of stream g, two samples behind lane
//     (g, j-1), passing the stereo sample down the lanes by shuffle.  Every band's recurrence stays strictly sequential
//     in the reference's operation order; left and right share the coefficients and ride in one packed f32x2 register
//     (FMUL2 / FFMA2 with a run-time 1.0 multiplier — ptxas 12.9 contracts mul.f32x2+add.f32x2 into FFMA2 even under
//     -fmad=false, which would break bit parity; an FFMA2 by an opaque 1.0 cannot be contracted and rounds once).
//     The warp also stages the next block's input rows into shared memory (cp.async, 16 B) while it filters.
//   * the other warps are the convolution warps, T = max(32, N/8) threads per stream: left + i*right go through ONE
//     complex N = 2B point Stockham FFT in shared memory (radix 8/4/2 in registers), the frequency-domain delay line
//     keeps that packed spectrum Z, and the four HRIR paths are applied as
//         W[k] = sum_p  Z_{t-p}[k] * A_p[k] + conj(Z_{t-p}[N-k]) * C_p[k]
//     with A = (G_L - i G_R)/2N, C = (G_L + i G_R)/2N, G_L = FFT(h_LSL + i h_LSR), G_R = FFT(h_RSL + i h_RSR):
//     Re IFFT(W) is the left ear (LSL + RSL), Im IFFT(W) the right ear (LSR + RSR) (src/dsp/convolution.rs:229-230).
//     One forward and one inverse FFT per block instead of the reference's four and four.
//   * the two roles are decoupled by named barriers over a 3-slot ring of filtered blocks, so the EQ of block t+1
//     overlaps the convolution of block t.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ohs {

constexpr int kMaxBands = 10;
constexpr int kEqGroup = 10;   // lanes per stream in the EQ warp
constexpr int kMaxG = 3;       // streams per CTA (3 x 10 lanes fill one warp)
constexpr int kEqSkew = 2;     // samples between neighbouring bands of the systolic chain (covers SHFL latency)
constexpr int kEqCoefStride = 8;  // floats per (eq_set, band): b0 b1 b2 a1 a2 enabled pad pad

enum NamedBarrier { kBarFull0 = 1, kBarFull1 = 2, kBarEmpty0 = 3, kBarEmpty1 = 4, kBarStream0 = 5 };

struct RenderParams {
    const float* in;            // [stream][2][row_stride]
    float* out;
    long long row_stride;       // frames
    int n_blocks;               // K engine blocks this launch
    int tail_frames;            // frames in the last block: B, or fewer in EQ-only mode (conv_enable == 0)
    int n_streams;
    const int* stream_hrir;     // [stream] -> hrir set
    const int* stream_eq;       // [stream] -> eq set
    const float* stream_gain;   // [stream]
    const float4* filt;         // [set][pmax][N] {A.re, A.im, C.re, C.im}, 1/N folded in
    const int* set_parts;       // [set] partitions in use
    float2* fdl;                // [stream][pmax][N] packed spectra ring (unused when every set has 1 partition)
    float2* prev;               // [stream][B] last filtered input block (overlap-save history)
    const float* eqc;           // [eq_set][kMaxBands][kEqCoefStride]
    float4* eqs;                // [stream][kMaxBands] {s1L, s1R, s2L, s2R}
    const float2* tw;           // [N] exp(-2*pi*i*m/N)
    int pmax;
    int head;                   // ring slot of this launch's first block
    int n_bands;
    int eq_enable;
    int conv_enable;
    float one;                  // 1.0f, deliberately opaque to the compiler (see header comment)
};

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// shared-memory index padding: one extra float2 every 8 keeps the stride-8/-64 accesses of the radix passes off the
// same 64-bit bank
__host__ __device__ constexpr int padi(int j) { return j + (j >> 3); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 3); }

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPending> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPending) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// register DFTs (forward, e^{-2 pi i rq/R}), natural order in and out
// ---------------------------------------------------------------------------------------------------------------
template <int R> struct Dft;
template <> struct Dft<2> {
    static __device__ __forceinline__ void run(float2 (&u)[2]) {
        const float2 t = u[0];
        u[0] = cadd(t, u[1]);
        u[1] = csub(t, u[1]);
    }
};
template <> struct Dft<4> {
    static __device__ __forceinline__ void run(float2 (&u)[4]) {
        const float2 a0 = cadd(u[0], u[2]), a1 = csub(u[0], u[2]);
        const float2 a2 = cadd(u[1], u[3]), a3 = mul_mi(csub(u[1], u[3]));
        u[0] = cadd(a0, a2); u[2] = csub(a0, a2);
        u[1] = cadd(a1, a3); u[3] = csub(a1, a3);
    }
};
template <> struct Dft<8> {
    static __device__ __forceinline__ void run(float2 (&u)[8]) {
        constexpr float c = 0.70710678118654752440f;
        float2 a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { a[r] = cadd(u[r], u[r + 4]); b[r] = csub(u[r], u[r + 4]); }
        // b[r] *= w8^r
        b[1] = make_float2(c * (b[1].x + b[1].y), c * (b[1].y - b[1].x));
        b[2] = mul_mi(b[2]);
        b[3] = make_float2(c * (b[3].y - b[3].x), -c * (b[3].x + b[3].y));
        Dft<4>::run(a);
        Dft<4>::run(b);
#pragma unroll
        for (int q = 0; q < 4; ++q) { u[2 * q] = a[q]; u[2 * q + 1] = b[q]; }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Stockham auto-sort FFT of N complex points by T threads (each thread owns E = N/T points per pass).
// Pass with radix R after radices of product P:  butterfly i -> k = i mod P, j = (i-k)*R + k,
//   u[r] = x[i + r*N/R] * w_N^{k r N/(P R)},  y[j + q*P] = DFT_R(u)[q].   Output in natural order.
// ---------------------------------------------------------------------------------------------------------------
template <int N, int T> struct FftPlan {
    static constexpr int E = N / T;
    static_assert(E == 8 || E == 4, "points per thread");
    static constexpr int R1 = E, P1 = 1;
    static constexpr int P2 = R1, R2 = (N / P2 >= E) ? E : N / P2;
    static constexpr int P3 = P2 * R2, R3 = (P3 < N) ? ((N / P3 >= E) ? E : N / P3) : 1;
    static constexpr int P4 = P3 * R3, R4 = (P4 < N) ? ((N / P4 >= E) ? E : N / P4) : 1;
    static_assert(P4 * R4 == N, "at most four passes");
    static constexpr int kPasses = 2 + (R3 > 1) + (R4 > 1);
    static constexpr bool kOutInB0 = (kPasses % 2) == 1;  // which ping-pong buffer a full transform ends in
};

template <int N, int T, int R, int P, class Load, class Store>
__device__ __forceinline__ void fft_pass(int tid, const float2* __restrict__ tw, Load load, Store store) {
    constexpr int NB = N / R;   // butterflies in this pass
    constexpr int PER = NB / T; // per thread
    static_assert(PER >= 1, "radix larger than points per thread");
#pragma unroll
    for (int b = 0; b < PER; ++b) {
        const int i = tid + b * T;
        const int k = i & (P - 1);
        const int j = (i - k) * R + k;
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = load(i + r * NB);
        if (P > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) u[r] = cmul(u[r], tw[(k * r) * (N / (P * R))]);
        }
        Dft<R>::run(u);
#pragma unroll
        for (int q = 0; q < R; ++q) store(j + q * P, u[q]);
    }
}

// Full transform.  load0 feeds the first pass, store_last receives the natural-order result, b0/b1 are the ping-pong
// buffers (padded), `sync` separates passes, `after_first` runs once the first pass has consumed its input.
template <int N, int T, class Load0, class StoreLast, class Sync, class AfterFirst>
__device__ __forceinline__ void fft_run(int tid, const float2* __restrict__ tw, float2* b0, float2* b1, Load0 load0,
                                        StoreLast store_last, Sync sync, AfterFirst after_first) {
    using Pl = FftPlan<N, T>;
    auto ld0 = [&](int i) { return b0[padi(i)]; };
    auto ld1 = [&](int i) { return b1[padi(i)]; };
    auto st0 = [&](int i, float2 v) { b0[padi(i)] = v; };
    auto st1 = [&](int i, float2 v) { b1[padi(i)] = v; };
    fft_pass<N, T, Pl::R1, Pl::P1>(tid, tw, load0, st0);
    after_first();
    sync();
    if constexpr (Pl::kPasses == 2) {
        fft_pass<N, T, Pl::R2, Pl::P2>(tid, tw, ld0, store_last);
    } else {
        fft_pass<N, T, Pl::R2, Pl::P2>(tid, tw, ld0, st1);
        sync();
        if constexpr (Pl::kPasses == 3) {
            fft_pass<N, T, Pl::R3, Pl::P3>(tid, tw, ld1, store_last);
        } else {
            fft_pass<N, T, Pl::R3, Pl::P3>(tid, tw, ld1, st0);
            sync();
            fft_pass<N, T, Pl::R4, Pl::P4>(tid, tw, ld0, store_last);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory carve-up of the render kernel
// ---------------------------------------------------------------------------------------------------------------
template <int N, int G> struct RenderSmem {
    static constexpr int B = N / 2;
    static constexpr int T = (N / 8 >= 32) ? N / 8 : 32;
    static constexpr int NP = padded_len(N);
    static constexpr int kThreads = 32 + G * T;
    static constexpr size_t kTwOff = 0;                                      // float2 tw[N]
    static constexpr size_t kZOff = kTwOff + sizeof(float2) * N;             // float2 z[G][2][NP]
    static constexpr size_t kRingOff = kZOff + sizeof(float2) * G * 2 * NP;  // float2 ring[G][3][B]
    static constexpr size_t kStageOff = kRingOff + sizeof(float2) * G * 3 * B;  // float stage[2][G][2][B]
    static constexpr size_t kBytes = kStageOff + sizeof(float) * 2 * G * 2 * B;
    // resident CTAs per SM the register allocation is held to: as many as shared memory and threads allow, up to four,
    // while leaving each thread at least 80 registers
    static constexpr int kBySmem = (int)((227 * 1024) / (kBytes + 1024));
    static constexpr int kByRegs = 65536 / (80 * kThreads);
    static constexpr int kMinBlocks0 = kBySmem < kByRegs ? kBySmem : kByRegs;
    static constexpr int kMinBlocks = kMinBlocks0 < 1 ? 1 : (kMinBlocks0 > 4 ? 4 : kMinBlocks0);
};

// ---------------------------------------------------------------------------------------------------------------
// EQ warp
// ---------------------------------------------------------------------------------------------------------------
// One DF2T step for the stereo pair, reference operation order (biquad 0.4.2 DirectForm2Transposed::run behind
// src/dsp/parametric_eq.rs:116-122):   out = s1 + b0*x;  s1 = (s2 + b1*x) - a1*out;  s2 = b2*x - a2*out
// Every product and every sum is rounded separately: x*c via FMUL2, sums via FFMA2(m, 1.0, s) == round(m + s).
__device__ __forceinline__ float2 df2t_step(float2 x, float2& s1, float2& s2, float2 b0, float2 b1, float2 b2, float2 na1,
                                            float2 na2, float2 one) {
    const float2 m0 = __fmul2_rn(b0, x);
    const float2 out = __ffma2_rn(m0, one, s1);
    const float2 m1 = __fmul2_rn(b1, x);
    const float2 t = __ffma2_rn(m1, one, s2);
    const float2 m2 = __fmul2_rn(na1, out);
    s1 = __ffma2_rn(m2, one, t);
    const float2 m3 = __fmul2_rn(b2, x);
    const float2 m4 = __fmul2_rn(na2, out);
    s2 = __ffma2_rn(m4, one, m3);
    return out;
}

// One engine block through the band-systolic chain of one warp.  Lane (g, j) filters sample n = step - D*j with band
// j; step runs 0 .. nb-1+9D.  The steady state (every lane busy) is branch-free and unrolled by four; the fill and
// drain steps (and ragged or partly disabled cases) go through the checked step, which commits state by select.
// FAST: nb == B and no valid lane is disabled.
template <int B, bool FAST>
__device__ __forceinline__ void eq_block_systolic(const float* __restrict__ xl, const float* __restrict__ xr, float2* __restrict__ dst,
                                                  int nb, int j, int src_lane, bool lane_valid, bool en, float2& s1, float2& s2,
                                                  float2 b0, float2 b1, float2 b2, float2 na1, float2 na2, float2 one) {
    static_assert(kEqSkew == 2, "register rotation below is written for a skew of two");
    [[maybe_unused]] constexpr int D = kEqSkew;
    constexpr int kLag = (kEqGroup - 1) * D;          // steps until the last band sees sample 0
    constexpr int kHead = (kLag + 3) / 4 * 4;         // checked steps before the unrolled steady state
    const bool first = (j == 0), last = (j == kEqGroup - 1) && lane_valid;  // lanes of absent streams never store
    float2 o0 = make_float2(0.f, 0.f), o1 = o0;       // this lane's outputs of the previous two steps (o1 older)

    auto checked_step = [&](int step) {
        const int n = step - D * j;
        const bool act = lane_valid && n >= 0 && n < nb;
        float2 x;
        x.x = __shfl_sync(0xffffffffu, o1.x, src_lane);
        x.y = __shfl_sync(0xffffffffu, o1.y, src_lane);
        const int nc = step < B ? step : B - 1;
        const float il = xl[nc], ir = xr[nc];
        x.x = first ? il : x.x;
        x.y = first ? ir : x.y;
        float2 t1 = s1, t2 = s2;
        float2 out = df2t_step(x, t1, t2, b0, b1, b2, na1, na2, one);
        const bool upd = act && en;
        s1.x = upd ? t1.x : s1.x; s1.y = upd ? t1.y : s1.y;
        s2.x = upd ? t2.x : s2.x; s2.y = upd ? t2.y : s2.y;
        out.x = upd ? out.x : x.x; out.y = upd ? out.y : x.y;
        if (act && last) dst[n] = out;
        o1 = o0; o0 = out;
    };

    if constexpr (!FAST) {
        for (int step = 0; step < nb + kLag; ++step) checked_step(step);
    } else {
#pragma unroll 1
    for (int step = 0; step < kHead; ++step) checked_step(step);
    // steady state: every lane holds a live sample; lanes of absent streams compute on garbage that is never stored
    float2* dlast = dst - kLag;
#pragma unroll 1
    for (int step = kHead; step < B; step += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(xl + step);
        const float4 r4 = *reinterpret_cast<const float4*>(xr + step);
        const float il[4] = {l4.x, l4.y, l4.z, l4.w};
        const float ir[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float2 x;
            x.x = __shfl_sync(0xffffffffu, o1.x, src_lane);
            x.y = __shfl_sync(0xffffffffu, o1.y, src_lane);
            x.x = first ? il[u] : x.x;
            x.y = first ? ir[u] : x.y;
            const float2 out = df2t_step(x, s1, s2, b0, b1, b2, na1, na2, one);
            if (last) dlast[step + u] = out;
            o1 = o0; o0 = out;
        }
    }
#pragma unroll 1
    for (int step = B; step < B + kLag; ++step) checked_step(step);
    }
}

template <int N, int G>
__device__ __forceinline__ void eq_warp_main(const RenderParams& p, unsigned char* smem, int stream0) {
    using SM = RenderSmem<N, G>;
    constexpr int B = SM::B;
    constexpr int D = kEqSkew;
    constexpr int kCount = SM::kThreads;
    float2* ring = reinterpret_cast<float2*>(smem + SM::kRingOff);
    float* stage = reinterpret_cast<float*>(smem + SM::kStageOff);

    const int lane = threadIdx.x;
    const int g_raw = lane / kEqGroup;
    const int j = lane - g_raw * kEqGroup;
    const int g = g_raw < G ? g_raw : G - 1;
    const int s = stream0 + g;
    const bool lane_valid = (g_raw < G) && (s < p.n_streams);
    const bool do_eq = p.eq_enable != 0;

    float2 b0 = make_float2(0.f, 0.f), b1 = b0, b2 = b0, na1 = b0, na2 = b0, s1 = b0, s2 = b0;
    bool en = false;
    const float2 one = make_float2(p.one, p.one);
    if (lane_valid && do_eq && j < p.n_bands) {
        const float* c = p.eqc + ((size_t)p.stream_eq[s] * kMaxBands + j) * kEqCoefStride;
        b0 = make_float2(c[0], c[0]); b1 = make_float2(c[1], c[1]); b2 = make_float2(c[2], c[2]);
        na1 = make_float2(-c[3], -c[3]); na2 = make_float2(-c[4], -c[4]);
        en = c[5] != 0.f;
        const float4 st = p.eqs[(size_t)s * kMaxBands + j];
        s1 = make_float2(st.x, st.y); s2 = make_float2(st.z, st.w);
    }

    // stage loader: rows (g', c) of block t -> stage[t&1][g'][c][0..B)
    auto issue_stage = [&](int t) {
        constexpr int kChunksPerRow = B / 4;
        constexpr int kChunks = G * 2 * kChunksPerRow;
        float* dst_base = stage + (size_t)(t & 1) * G * 2 * B;
        const int nb = (t == p.n_blocks - 1) ? p.tail_frames : B;
        if (nb == B) {
            for (int q = lane; q < kChunks; q += 32) {
                const int row = q / kChunksPerRow, off = q - row * kChunksPerRow;
                const int sg = stream0 + (row >> 1);
                if (sg < p.n_streams) {
                    const float* src = p.in + ((size_t)sg * 2 + (row & 1)) * p.row_stride + (size_t)t * B + off * 4;
                    cp_async16(dst_base + row * B + off * 4, src);
                }
            }
        } else {
            // ragged last block (EQ-only mode, any host-buffer length): plain guarded loads
            for (int q = lane; q < G * 2 * B; q += 32) {
                const int row = q / B, n = q - row * B;
                const int sg = stream0 + (row >> 1);
                if (sg < p.n_streams && n < nb)
                    dst_base[row * B + n] = p.in[((size_t)sg * 2 + (row & 1)) * p.row_stride + (size_t)t * B + n];
            }
        }
        cp_async_commit();
    };

    const int src_lane = (j == 0) ? lane : lane - 1;
    // every valid lane filters (no disabled band in this warp): the steady-state loop needs no per-lane selects
    const bool all_fast = __all_sync(0xffffffffu, en || !lane_valid);
    issue_stage(0);
    for (int t = 0; t < p.n_blocks; ++t) {
        if (t + 1 < p.n_blocks) { issue_stage(t + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncwarp();
        if (t >= 2) bar_sync(kBarEmpty0 + (t & 1), kCount);  // ring slot t%3 was last read as history of block t-2
        const int slot = t % 3;
        const int nb = (t == p.n_blocks - 1) ? p.tail_frames : B;
        const float* st_base = stage + (size_t)(t & 1) * G * 2 * B;
        if (!do_eq) {
            // EQ off (src/lib.rs:1179): the warp only interleaves left/right into the ring
            for (int q = lane; q < G * B; q += 32) {
                const int gg = q / B, n = q - gg * B;
                ring[(gg * 3 + slot) * B + n] = make_float2(st_base[(gg * 2) * B + n], st_base[(gg * 2 + 1) * B + n]);
            }
        } else {
            const float* xl = st_base + (g * 2) * B;
            const float* xr = xl + B;
            float2* dst = ring + (g * 3 + slot) * B;
            if (nb == B && all_fast) eq_block_systolic<B, true>(xl, xr, dst, B, j, src_lane, lane_valid, en, s1, s2, b0, b1, b2, na1, na2, one);
            else eq_block_systolic<B, false>(xl, xr, dst, nb, j, src_lane, lane_valid, en, s1, s2, b0, b1, b2, na1, na2, one);
        }
        __threadfence_block();
        bar_arrive(kBarFull0 + (t & 1), kCount);
    }
    if (lane_valid && do_eq && j < p.n_bands) p.eqs[(size_t)s * kMaxBands + j] = make_float4(s1.x, s1.y, s2.x, s2.y);
}

// ---------------------------------------------------------------------------------------------------------------
// convolution warps
// ---------------------------------------------------------------------------------------------------------------
// W[ka] += u*A[ka] + conj(pu)*C[ka]
__device__ __forceinline__ void mac_bin(float2& acc, float2 u, float2 pu, float4 f) {
    acc.x = fmaf(u.x, f.x, acc.x); acc.x = fmaf(-u.y, f.y, acc.x);
    acc.y = fmaf(u.x, f.y, acc.y); acc.y = fmaf(u.y, f.x, acc.y);
    acc.x = fmaf(pu.x, f.z, acc.x); acc.x = fmaf(pu.y, f.w, acc.x);
    acc.y = fmaf(pu.x, f.w, acc.y); acc.y = fmaf(-pu.y, f.z, acc.y);
}

template <int N, int G>
__device__ __forceinline__ void conv_warps_main(const RenderParams& p, unsigned char* smem, int stream0) {
    using SM = RenderSmem<N, G>;
    constexpr int B = SM::B, T = SM::T, NP = SM::NP;
    constexpr int kCount = SM::kThreads;
    using Pl = FftPlan<N, T>;
    const float2* tw = reinterpret_cast<const float2*>(smem + SM::kTwOff);
    float2* ring = reinterpret_cast<float2*>(smem + SM::kRingOff);

    const int ft = threadIdx.x - 32;
    const int g = ft / T, tid = ft - g * T;
    const int s = stream0 + g;
    const bool valid = s < p.n_streams;
    float2* b0 = reinterpret_cast<float2*>(smem + SM::kZOff) + (size_t)g * 2 * NP;
    float2* b1 = b0 + NP;
    float2* zbuf = Pl::kOutInB0 ? b0 : b1;   // forward transform lands here
    float2* wbuf = Pl::kOutInB0 ? b1 : b0;   // frequency-domain product goes here
    float2* ring_g = ring + (size_t)g * 3 * B;

    int nparts = 1;
    const float4* filt = p.filt;
    float gain = 1.f;
    float2* fdl_s = nullptr;
    if (valid) {
        const int set = p.stream_hrir[s];
        nparts = p.set_parts[set];
        filt = p.filt + (size_t)set * p.pmax * N;
        gain = p.stream_gain[s];
        fdl_s = p.fdl + (size_t)s * p.pmax * N;
    }
    float* out_l = p.out + ((size_t)s * 2) * p.row_stride;
    float* out_r = out_l + p.row_stride;
    auto stream_sync = [&]() { if (T > 32) bar_sync(kBarStream0 + g, T); else __syncwarp(); };

    for (int t = 0; t < p.n_blocks; ++t) {
        bar_sync(kBarFull0 + (t & 1), kCount);
        const int cur = t % 3, prv = (t + 2) % 3;
        const float2* xc = ring_g + cur * B;
        const float2* xp = ring_g + prv * B;
        const bool release = (t + 2 < p.n_blocks);
        if (!valid) {
            if (release) bar_arrive(kBarEmpty0 + (t & 1), kCount);
            continue;
        }
        if (!p.conv_enable) {
            // EQ + gain only (StereoParametricEQ::process_block followed by the gain loop)
            const int nb = (t == p.n_blocks - 1) ? p.tail_frames : B;
            for (int n = tid; n < nb; n += T) {
                const float2 v = xc[n];
                out_l[(size_t)t * B + n] = v.x * gain;
                out_r[(size_t)t * B + n] = v.y * gain;
            }
            if (release) bar_arrive(kBarEmpty0 + (t & 1), kCount);
            continue;
        }
        // ---- forward FFT of the overlap-save window [previous block | current block], z = left + i*right
        fft_run<N, T>(
            tid, tw, b0, b1, [&](int i) { return i < B ? xp[i] : xc[i - B]; },
            [&](int i, float2 v) { zbuf[padi(i)] = v; }, stream_sync,
            [&]() { if (release) bar_arrive(kBarEmpty0 + (t & 1), kCount); });
        stream_sync();
        // ---- frequency-domain delay line + 4-path multiply-accumulate
        int slot = p.head + t;
        slot -= (slot / p.pmax) * p.pmax;
        if (nparts > 1) {
            float2* dstz = fdl_s + (size_t)slot * N;
#pragma unroll
            for (int e = 0; e < N / T; ++e) { const int i = tid + e * T; dstz[i] = zbuf[padi(i)]; }
        }
        constexpr int kItems = N / 2 / T;
        float2 acc_a[kItems], acc_b[kItems];
#pragma unroll
        for (int m = 0; m < kItems; ++m) {
            const int k = tid + m * T;
            const int ka = k, kb = k ? N - k : N / 2;
            const float2 u = zbuf[padi(ka)], v = zbuf[padi(kb)];
            const float4 fa = __ldg(filt + ka), fb = __ldg(filt + kb);
            acc_a[m] = make_float2(0.f, 0.f); acc_b[m] = make_float2(0.f, 0.f);
            mac_bin(acc_a[m], u, k ? v : u, fa);
            mac_bin(acc_b[m], v, k ? u : v, fb);
        }
        for (int q = 1; q < nparts; ++q) {
            int sl = slot - q; if (sl < 0) sl += p.pmax;
            const float2* zq = fdl_s + (size_t)sl * N;
            const float4* fq = filt + (size_t)q * N;
#pragma unroll
            for (int m = 0; m < kItems; ++m) {
                const int k = tid + m * T;
                const int ka = k, kb = k ? N - k : N / 2;
                const float2 u = zq[ka], v = zq[kb];
                const float4 fa = __ldg(fq + ka), fb = __ldg(fq + kb);
                mac_bin(acc_a[m], u, k ? v : u, fa);
                mac_bin(acc_b[m], v, k ? u : v, fb);
            }
        }
#pragma unroll
        for (int m = 0; m < kItems; ++m) {
            const int k = tid + m * T;
            const int ka = k, kb = k ? N - k : N / 2;
            // swap(re, im): the inverse transform is run as swap(FFT(swap(W)))
            wbuf[padi(ka)] = make_float2(acc_a[m].y, acc_a[m].x);
            wbuf[padi(kb)] = make_float2(acc_b[m].y, acc_b[m].x);
        }
        stream_sync();
        // ---- inverse FFT; keep the last B samples (overlap-save), ear sums are already inside W, apply gain
        float* ol = out_l + (size_t)t * B - B;
        float* orr = out_r + (size_t)t * B - B;
        fft_run<N, T>(
            tid, tw, zbuf, wbuf, [&](int i) { return wbuf[padi(i)]; },
            [&](int i, float2 v) { if (i >= B) { ol[i] = v.y * gain; orr[i] = v.x * gain; } }, stream_sync, [&]() {});
    }
    // overlap-save history for the next launch: the last filtered block
    if (valid && p.conv_enable && p.n_blocks > 0) {
        const float2* xc = ring_g + ((p.n_blocks - 1) % 3) * B;
        for (int n = tid; n < B; n += T) p.prev[(size_t)s * B + n] = xc[n];
    }
}

template <int N, int G>
__global__ void __launch_bounds__(RenderSmem<N, G>::kThreads, RenderSmem<N, G>::kMinBlocks) render_kernel(const RenderParams p) {
    using SM = RenderSmem<N, G>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int stream0 = blockIdx.x * G;
    {
        float2* tw = reinterpret_cast<float2*>(smem + SM::kTwOff);
        for (int i = threadIdx.x; i < N; i += SM::kThreads) tw[i] = p.tw[i];
        // overlap-save history -> ring slot 2 (the "previous" slot of block 0)
        float2* ring = reinterpret_cast<float2*>(smem + SM::kRingOff);
        for (int q = threadIdx.x; q < G * SM::B; q += SM::kThreads) {
            const int g = q / SM::B, n = q - g * SM::B;
            const int s = stream0 + g;
            ring[(g * 3 + 2) * SM::B + n] = (s < p.n_streams) ? p.prev[(size_t)s * SM::B + n] : make_float2(0.f, 0.f);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) eq_warp_main<N, G>(p, smem, stream0);
    else conv_warps_main<N, G>(p, smem, stream0);
}

// ---------------------------------------------------------------------------------------------------------------
// HRIR set-up: ConvolutionEngine::set_ir (src/dsp/convolution.rs:111-139) for one (set, partition) per CTA.
// ir: [set][4][pmax*B] zero-padded time-domain taps; filt: [set][pmax][N]
// ---------------------------------------------------------------------------------------------------------------
template <int N> struct SetupSmem {
    static constexpr int T = (N / 8 >= 32) ? N / 8 : 32;
    static constexpr int NP = padded_len(N);
    static constexpr size_t kBytes = sizeof(float2) * (N + 4 * NP);
};

template <int N>
__global__ void __launch_bounds__(SetupSmem<N>::T) setup_filters_kernel(const float* __restrict__ ir, float4* __restrict__ filt,
                                                                      const float2* __restrict__ tw_g,
                                                                      const int* __restrict__ set_list,
                                                                      const int* __restrict__ set_parts, int pmax) {
    using SM = SetupSmem<N>;
    constexpr int B = N / 2, T = SM::T, NP = SM::NP;
    using Pl = FftPlan<N, T>;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* tw = reinterpret_cast<float2*>(smem);
    float2* a0 = tw + N;
    float2* a1 = a0 + NP;
    float2* c0 = a1 + NP;
    float2* c1 = c0 + NP;
    const int tid = threadIdx.x;
    const int part = blockIdx.x;
    const int set = set_list[blockIdx.y];
    if (part >= set_parts[set]) return;
    for (int i = tid; i < N; i += T) tw[i] = tw_g[i];
    __syncthreads();
    const float* h = ir + (size_t)set * 4 * pmax * B + (size_t)part * B;
    const size_t ps = (size_t)pmax * B;  // path stride
    auto sync = [&]() { __syncthreads(); };
    float2* gl = Pl::kOutInB0 ? a0 : a1;
    float2* gr = Pl::kOutInB0 ? c0 : c1;
    // G_L = FFT(h_LSL + i h_LSR), G_R = FFT(h_RSL + i h_RSR), each chunk zero-padded to N (:123-129)
    fft_run<N, T>(
        tid, tw, a0, a1, [&](int i) { return i < B ? make_float2(h[i], h[ps + i]) : make_float2(0.f, 0.f); },
        [&](int i, float2 v) { gl[padi(i)] = v; }, sync, [&]() {});
    fft_run<N, T>(
        tid, tw, c0, c1, [&](int i) { return i < B ? make_float2(h[2 * ps + i], h[3 * ps + i]) : make_float2(0.f, 0.f); },
        [&](int i, float2 v) { gr[padi(i)] = v; }, sync, [&]() {});
    __syncthreads();
    const float sc = 1.0f / (2.0f * (float)N);  // 1/2 of the real/imag split and the 1/FFT_SIZE of :280, exact power of two
    float4* dst = filt + ((size_t)set * pmax + part) * N;
    for (int k = tid; k < N; k += T) {
        const float2 l = gl[padi(k)], r = gr[padi(k)];
        dst[k] = make_float4((l.x + r.y) * sc, (l.y - r.x) * sc, (l.x - r.y) * sc, (l.y + r.x) * sc);
    }
}

// zero a stream's convolution history (delay line + overlap-save block) for streams bound to a flagged set, or all
// streams when set_flags is null
__global__ void clear_history_kernel(float2* fdl, float2* prev, const int* stream_hrir, int n_streams,
                                     const unsigned char* set_flags, size_t fdl_per_stream, size_t prev_per_stream) {
    const int s = blockIdx.x;
    if (s >= n_streams) return;
    if (set_flags && !set_flags[stream_hrir[s]]) return;
    float2* f = fdl + (size_t)s * fdl_per_stream;
    for (size_t i = threadIdx.x; i < fdl_per_stream; i += blockDim.x) f[i] = make_float2(0.f, 0.f);
    float2* q = prev + (size_t)s * prev_per_stream;
    for (size_t i = threadIdx.x; i < prev_per_stream; i += blockDim.x) q[i] = make_float2(0.f, 0.f);
}

}  // namespace ohs
