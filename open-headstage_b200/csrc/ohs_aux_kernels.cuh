// ohs_aux_kernels.cuh — the kernels around the fused render kernel: HRIR set-up (set_ir), the time-batched long-response
// route (history gather, per-bin convolution along time, inverse transforms), history reset and the object mixdown.
// Included by ohs_api.cu only (the non-template kernels here have external linkage); the render kernel's templates
// live in ohs_kernels.cuh and are instantiated per transform size in ohs_render.cu.
#pragma once

#include "ohs_kernels.cuh"

namespace ohs {

// Threads per transform of the stand-alone forward and inverse kernels: twice the render kernel's where that keeps the
// radix plan (8 points per thread instead of 16: same butterflies, same twiddles, bit-identical spectra), because these
// kernels are bound by memory latency and want warps, not registers
constexpr int xform_threads(int n) { return (n / fft_threads(n) >= 16) ? 2 * fft_threads(n) : fft_threads(n); }

// ---------------------------------------------------------------------------------------------------------------
// HRIR set-up: ConvolutionEngine::set_ir (src/dsp/convolution.rs:111-139) for one (set, partition) per CTA.
// ir: [set][4][pmax*B] zero-padded time-domain taps; filt: [set][pmax][N]
// ---------------------------------------------------------------------------------------------------------------
template <int N> struct SetupSmem {
    static constexpr int T = fft_threads(N);
    static constexpr int TX = xform_threads(N);   // the stand-alone forward and inverse kernels of the time-batched route
    static_assert(FftPlan<N, TX>::R1 == FftPlan<N, T>::R1 && FftPlan<N, TX>::R2 == FftPlan<N, T>::R2 && FftPlan<N, TX>::R3 == FftPlan<N, T>::R3 &&
                  FftPlan<N, TX>::R4 == FftPlan<N, T>::R4, "same radix plan, hence bit-identical transforms");
    static constexpr int NP = padded_len(N);
    static constexpr size_t kBytes = sizeof(float2) * (N + 4 * NP);
};

// first-pass loader: an impulse-response chunk of two paths as (re, im), zero-padded from B to N (:123-129)
struct IrChunk {
    const float* ha; const float* hb; int B;
    __device__ __forceinline__ float2 ld(int i) const { return i < B ? make_float2(ha[i], hb[i]) : make_float2(0.f, 0.f); }
    __device__ __forceinline__ void ld2(int i, float2& a, float2& b) const { a = ld(i); b = ld(i + 1); }
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
    const SmemCx gl{Pl::kOutInB0 ? a0 : a1}, gr{Pl::kOutInB0 ? c0 : c1};
    // G_L = FFT(h_LSL + i h_LSR), G_R = FFT(h_RSL + i h_RSR)
    fft_run<N, T>(tid, tw, a0, a1, IrChunk{h, h + ps, B}, gl, sync, [&]() {});
    fft_run<N, T>(tid, tw, c0, c1, IrChunk{h + 2 * ps, h + 3 * ps, B}, gr, sync, [&]() {});
    __syncthreads();
    const float sc = 1.0f / (2.0f * (float)N);  // 1/2 of the real/imag split and the 1/FFT_SIZE of :280, exact power of two
    float4* dst = filt + ((size_t)set * pmax + part) * N;
    for (int k = tid; k < N; k += T) {
        const float2 l = gl.ld(k), r = gr.ld(k);
        // even bins first, odd bins behind them: the layout the convolution warps' lane-contiguous loads want
        dst[(k & 1) * (N / 2) + (k >> 1)] = make_float4((l.x + r.y) * sc, (l.y - r.x) * sc, (l.x - r.y) * sc, (l.y + r.x) * sc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Time-batched long responses (the offline per-bin pipeline).  The fused kernel re-reads every stream's
// delay line (P-1 spectra of 8N bytes) for every block; for a launch of K blocks with a long response that is K(P-1)
// spectrum reads per stream where K + P - 1 would do.  In this mode the render kernel only filters and transforms
// (EQ, forward FFT, spectra into the delay-line ring AND into a buffer in time order), then per bin
//     W_t[k] = sum_q  Z_{t-q}[k] A_q[k] + conj(Z_{t-q}[N-k]) C_q[k]
// is a short convolution ALONG TIME: a thread owns a bin couple (k, N-k) of one stream and TB consecutive blocks,
// keeps their TB accumulators and a sliding window of TB spectra in registers, and walks q once — one new spectrum
// value and one filter tap loaded per step for TB*16 FMAs.  A third kernel runs the inverse transforms.
// ---------------------------------------------------------------------------------------------------------------
// The time-ordered buffer is CIRCULAR in time: `cap` slots per stream, the slot of the next block (`zbase`) moves on
// by the blocks rendered, and the pmax-1 slots behind it hold the history — so consecutive time-batched sub-launches
// and calls find their history in place and nothing is copied between them.  The delay-line ring (what the
// block-by-block kernel reads) and the buffer are converted into each other only when the route changes:
// ring -> buffer (time -q, q = 1..pmax-1, goes to slot zbase - q) ...
__global__ void gather_history_kernel(const float2* __restrict__ fdl, float2* __restrict__ zlin, int N, int pmax, int head,
                                      long long zlin_stride, int zbase, int cap) {
    const int q = blockIdx.x + 1, s = blockIdx.y;
    int sl = head - q; if (sl < 0) sl += pmax;
    int zs = zbase - q; if (zs < 0) zs += cap;
    const float4* src = reinterpret_cast<const float4*>(fdl + ((size_t)s * pmax + sl) * N);
    float4* dst = reinterpret_cast<float4*>(zlin + (size_t)s * zlin_stride + (size_t)zs * N);
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) dst[i] = src[i];
}
// ... and buffer -> ring
__global__ void scatter_history_kernel(const float2* __restrict__ zlin, float2* __restrict__ fdl, int N, int pmax, int head,
                                       long long zlin_stride, int zbase, int cap) {
    const int q = blockIdx.x + 1, s = blockIdx.y;
    int sl = head - q; if (sl < 0) sl += pmax;
    int zs = zbase - q; if (zs < 0) zs += cap;
    const float4* src = reinterpret_cast<const float4*>(zlin + (size_t)s * zlin_stride + (size_t)zs * N);
    float4* dst = reinterpret_cast<float4*>(fdl + ((size_t)s * pmax + sl) * N);
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) dst[i] = src[i];
}

// Block = 4 warps = 4 streams x 32 bin couples: the four warps read the same filter taps (they hit in L1), each its own
// stream's spectra (32 couples x 8 bytes contiguous per load).  The window of TB spectra lives in shared memory, one
// 16-byte column entry per thread and slot (slot = time mod TB), so the step loop needs no unrolling over the
// window's rotation: fully unrolled it is 90 KB of code and the kernel stalls on instruction fetch.
// Two steps (partitions q and q+1) per iteration: the window value of time tau meets tap q in block tau+q and tap q+1 in
// block tau+q+1, so every value read from the shared-memory window feeds 32 FMAs instead of 16 — the per-step form
// spent as many shared-memory wavefront cycles as FMA issue cycles (16 LDS.128 per 256 FMAs per thread).  The value of
// time t0-q-1, which block t0 needs for tap q+1, is the one that has just been loaded from the time-ordered buffer and
// is still in registers.  An odd last partition runs the single step.
template <int TB, bool kDc>
__device__ __forceinline__ void bin_conv_steps(float2 (&a0)[TB], float2 (&a1)[TB], float4* win, const float4* __restrict__ f,
                                               const float2* __restrict__ z, int N, int nparts, int zslot, int cap, int f0i, int f1i,
                                               int k0, int k1) {
    // One iteration's operands: the taps of partitions q and q+1 for the thread's two bins, and the spectra of times
    // t0-q-1 and t0-q-2, which enter the window behind them.  They are loaded one iteration ahead (an iteration is ~570
    // instructions, more than a loaded HBM round trip) into the register set the other iteration does not use (the
    // loop is unrolled by two by hand: no register moves), taps through running pointers, spectra through a running slot.
    struct Ops { float4 fa0, fb0, fa1, fb1, n0, n1; };
    // `z`: the stream's circular time-ordered buffer (cap slots); `zslot`: the slot of the thread's first block t0
    const float4* pf0 = f + f0i;
    const float4* pf1 = f + f1i;
    const long long N2 = 2ll * N;
    int sz = zslot - 1; if (sz < 0) sz += cap;   // slot of time t0-q-1, walking back two slots per iteration
    auto load = [&](Ops& o, int q) {
        if (q < nparts) { o.fa0 = pf0[0]; o.fb0 = pf1[0]; }
        if (q + 1 < nparts) { o.fa1 = pf0[N]; o.fb1 = pf1[N]; }
        // the spectra loads are unconditional: a wrapped slot is always inside the stream's buffer, and a value that no
        // partition needs is never used (a conditional load here compiles to a branch that pins the load behind it)
        int s1 = sz - 1; if (s1 < 0) s1 += cap;
        const float2* zp0 = z + (long long)sz * N;
        const float2* zp1 = z + (long long)s1 * N;
        const float2 x0 = zp0[k0], y0 = zp0[k1], x1 = zp1[k0], y1 = zp1[k1];
        o.n0 = make_float4(x0.x, x0.y, y0.x, y0.y);
        o.n1 = make_float4(x1.x, x1.y, y1.x, y1.y);
        pf0 += N2; pf1 += N2;
        sz = s1 - 1; if (sz < 0) sz += cap;
    };
    auto pair_step = [&](const Ops& o, int q) {
        const int rot = (TB - (q & (TB - 1))) & (TB - 1);   // slot of time t0+tb-q is (tb + rot) mod TB
        const float4* wrot = win + rot * 128;               // every slot is stored twice, TB slots apart: no wrap in the reads
        {
            const float2 u0 = make_float2(o.n0.x, o.n0.y), u1 = make_float2(o.n0.z, o.n0.w);   // time t0-q-1, tap q+1, block t0
            mac_bin(a0[0], u0, kDc ? u0 : u1, o.fa1);
            mac_bin(a1[0], u1, kDc ? u1 : u0, o.fb1);
        }
#pragma unroll
        for (int tb = 0; tb < TB; ++tb) {
            const float4 v = wrot[tb * 128];
            const float2 u0 = make_float2(v.x, v.y), u1 = make_float2(v.z, v.w);
            mac_bin(a0[tb], u0, kDc ? u0 : u1, o.fa0);   // couple 0 = the two self-mirrored bins 0 and N/2
            mac_bin(a1[tb], u1, kDc ? u1 : u0, o.fb0);
            if (tb + 1 < TB) {
                mac_bin(a0[tb + 1], u0, kDc ? u0 : u1, o.fa1);
                mac_bin(a1[tb + 1], u1, kDc ? u1 : u0, o.fb1);
            }
        }
        // the window moves two blocks back: times t0-q-1 and t0-q-2 replace times t0+TB-1-q and t0+TB-2-q
        const int e1 = (TB - 1 + rot) & (TB - 1), e2 = (TB - 2 + rot) & (TB - 1);
        win[e1 * 128] = o.n0; win[(e1 + TB) * 128] = o.n0;
        win[e2 * 128] = o.n1; win[(e2 + TB) * 128] = o.n1;
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    Ops A{zero4, zero4, zero4, zero4, zero4, zero4}, B = A;
    load(A, 0);
    const int npairs = nparts >> 1;
#pragma unroll 1
    for (int p = 0; p < npairs; p += 2) {
        load(B, 2 * p + 2);
        pair_step(A, 2 * p);
        if (p + 1 < npairs) {
            load(A, 2 * p + 4);
            pair_step(B, 2 * p + 2);
        }
    }
    if (nparts & 1) {   // odd partition count: one single step (no later step reads the window)
        const int q = nparts - 1;
        const bool in_a = (npairs & 1) == 0;   // the set the last load went to
        const float4 fa = in_a ? A.fa0 : B.fa0, fb = in_a ? A.fb0 : B.fb0;
        const int rot = (TB - (q & (TB - 1))) & (TB - 1);
        const float4* wrot = win + rot * 128;
#pragma unroll
        for (int tb = 0; tb < TB; ++tb) {
            const float4 v = wrot[tb * 128];
            const float2 u0 = make_float2(v.x, v.y), u1 = make_float2(v.z, v.w);
            mac_bin(a0[tb], u0, kDc ? u0 : u1, fa);
            mac_bin(a1[tb], u1, kDc ? u1 : u0, fb);
        }
    }
}

template <int TB>
__global__ void __launch_bounds__(128, 3) bin_conv_kernel(const float2* __restrict__ zlin, float2* __restrict__ wlin,
                                                       const float4* __restrict__ filt, const int* __restrict__ stream_hrir,
                                                       const int* __restrict__ set_parts, int N, int pmax, int K, int n_streams,
                                                       long long zlin_stride, int zslot0, int cap) {
    static_assert((TB & (TB - 1)) == 0, "window slots are indexed modulo a power of two");
    extern __shared__ __align__(16) unsigned char smem[];
    float4* win = reinterpret_cast<float4*>(smem) + threadIdx.x;   // this thread's column: slot e at win[e * 128]
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);   // couple index: bins (i, N-i); couple 0 is (0, N/2)
    const int t0 = blockIdx.y * TB, s = blockIdx.z * 4 + (threadIdx.x >> 5);
    if (i >= N / 2 || s >= n_streams) return;
    const int k0 = i, k1 = i ? N - i : N / 2;
    const int set = stream_hrir[s];
    const int nparts = set_parts[set];
    const float4* f = filt + (size_t)set * pmax * N;
    const int f0i = (k0 & 1) * (N / 2) + (k0 >> 1), f1i = (k1 & 1) * (N / 2) + (k1 >> 1);   // even-bins-first table layout
    const float2* z = zlin + (size_t)s * zlin_stride;   // circular: block t of this launch sits in slot (zslot0 + t) mod cap
    int zs = zslot0 + t0; if (zs >= cap) zs -= cap;
    float2 a0[TB], a1[TB];
    {
        // the window's first TB spectra: every load unconditional (blocks past the launch's last re-read that one and are
        // zeroed by a select) and issued before the first store, so that the TB round trips overlap
        float2 v0[TB], v1[TB];
        const int last = K - 1 - t0;
#pragma unroll
        for (int tb = 0; tb < TB; ++tb) {
            int sl = zs + (tb < last ? tb : last); if (sl >= cap) sl -= cap;
            v0[tb] = z[(size_t)sl * N + k0];
            v1[tb] = z[(size_t)sl * N + k1];
        }
#pragma unroll
        for (int tb = 0; tb < TB; ++tb) {
            a0[tb] = make_float2(0.f, 0.f); a1[tb] = make_float2(0.f, 0.f);
            const bool live = tb <= last;
            const float4 v = live ? make_float4(v0[tb].x, v0[tb].y, v1[tb].x, v1[tb].y) : make_float4(0.f, 0.f, 0.f, 0.f);
            win[tb * 128] = v;
            win[(tb + TB) * 128] = v;
        }
    }
    // couple 0 (bins 0 and N/2, each its own mirror) lives in lane 0 of the first couple group's warps only
    if (blockIdx.x == 0) {
        if (i == 0) bin_conv_steps<TB, true>(a0, a1, win, f, z, N, nparts, zs, cap, f0i, f1i, k0, k1);
        else bin_conv_steps<TB, false>(a0, a1, win, f, z, N, nparts, zs, cap, f0i, f1i, k0, k1);
    } else {
        bin_conv_steps<TB, false>(a0, a1, win, f, z, N, nparts, zs, cap, f0i, f1i, k0, k1);
    }
    float2* w = wlin + ((size_t)s * K) * N;
#pragma unroll
    for (int tb = 0; tb < TB; ++tb)
        if (t0 + tb < K) { w[(size_t)(t0 + tb) * N + k0] = a0[tb]; w[(size_t)(t0 + tb) * N + k1] = a1[tb]; }
}

// first-pass loader of the forward transform from global memory: the overlap-save window [previous block | current block]
// of a stream's EQ-filtered rows, z = left + i*right
struct RowWindow {
    const float* pl; const float* pr; const float* cl; const float* cr; int B;   // cl/cr already offset by -B
    __device__ __forceinline__ float2 ld(int i) const { return i < B ? make_float2(pl[i], pr[i]) : make_float2(cl[i], cr[i]); }
    __device__ __forceinline__ void ld2(int i, float2& a, float2& b) const {  // i even: samples i and i+1
        const float2 l = *reinterpret_cast<const float2*>((i < B ? pl : cl) + i), r = *reinterpret_cast<const float2*>((i < B ? pr : cr) + i);
        a = make_float2(l.x, r.x); b = make_float2(l.y, r.y);
    }
};

// last-pass store of the forward transform: the packed spectrum goes straight from registers to its slot of the
// time-ordered buffer
struct SpectrumStore {
    float2* zl;
    __device__ __forceinline__ void st(int i, float2 v) const { zl[i] = v; }
    __device__ __forceinline__ void st2(int i, float2 a, float2 b) const {
        *reinterpret_cast<float4*>(zl + i) = make_float4(a.x, a.y, b.x, b.y);
    }
};

// one CTA per (block, stream): forward transform of block t0 + blockIdx.x of the filtered rows `xf` (block 0's history is
// the engine's overlap-save block `prev`), same plan and rounding as the render kernel's forward transform, into slot
// (zlin_base + t) mod cap of the circular time-ordered buffer.  (The delay-line ring is not written here: it is rebuilt
// from the buffer when something needs it, scatter_history_kernel.)
template <int N>
__global__ void __launch_bounds__(SetupSmem<N>::TX) forward_kernel(const float* __restrict__ xf, long long xf_stride, int t0,
                                                                 const float* __restrict__ prev, float2* __restrict__ zlin,
                                                                 long long zlin_stride, int zlin_base, int cap,
                                                                 const float2* __restrict__ tw_g) {
    constexpr int T = SetupSmem<N>::TX, NP = SetupSmem<N>::NP, B = N / 2;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* b0 = reinterpret_cast<float2*>(smem);
    float2* b1 = b0 + NP;
    const int t = t0 + blockIdx.x, s = blockIdx.y;
    const float* cl = xf + ((size_t)s * 2) * xf_stride + (size_t)t * B;
    const float* cr = cl + xf_stride;
    const float* pl = t ? cl - B : prev + (size_t)s * 2 * B;
    const float* pr = t ? cr - B : pl + B;
    int zs = zlin_base + t; if (zs >= cap) zs -= cap;
    float2* zl = zlin + (size_t)s * zlin_stride + (size_t)zs * N;
    auto sync = [&]() { __syncthreads(); };
    fft_run<N, T>(threadIdx.x, tw_g, b0, b1, RowWindow{pl, pr, cl - B, cr - B, B}, SpectrumStore{zl}, sync, [&]() {});
}

// first-pass loader of the inverse transform from global memory, with the swap of swap o FFT o swap
struct SpectrumSwapLoad {
    const float2* w;
    __device__ __forceinline__ float2 ld(int i) const { const float2 v = w[i]; return make_float2(v.y, v.x); }
    __device__ __forceinline__ void ld2(int i, float2& a, float2& b) const {
        const float4 v = *reinterpret_cast<const float4*>(w + i);
        a = make_float2(v.y, v.x); b = make_float2(v.w, v.z);
    }
};

// one CTA per (block, stream): inverse transform of W_t, last B samples times gain to the output rows
template <int N>
__global__ void __launch_bounds__(SetupSmem<N>::TX) inverse_kernel(const float2* __restrict__ wlin, float* __restrict__ out,
                                                                const float2* __restrict__ tw_g, const float* __restrict__ stream_gain,
                                                                int K, long long row_stride) {
    constexpr int T = SetupSmem<N>::TX, NP = SetupSmem<N>::NP, B = N / 2;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* b0 = reinterpret_cast<float2*>(smem);
    float2* b1 = b0 + NP;
    const int t = blockIdx.x, s = blockIdx.y;
    const float2* w = wlin + ((size_t)s * K + t) * N;
    float* out_l = out + ((size_t)s * 2) * row_stride;
    float* out_r = out_l + row_stride;
    auto sync = [&]() { __syncthreads(); };
    fft_run<N, T>(threadIdx.x, tw_g, b0, b1, SpectrumSwapLoad{w},
                  OutputStore{out_l + (size_t)t * B - B, out_r + (size_t)t * B - B, stream_gain[s], B}, sync, [&]() {});
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

// Object mixdown (BASELINE config 4): bus[c][n] = sum over streams of in[s][c][n].  Sequential f32 sum in stream order
// (deterministic); one thread per four frames, rows read with coalesced 16-byte loads.
__global__ void mix_streams_kernel(const float* __restrict__ in, float* __restrict__ bus, int n_streams, size_t n_frames,
                                   size_t row_stride, size_t bus_stride) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index inside a row
    const int c = blockIdx.y;
    if (q * 4 >= n_frames) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* row = in + (size_t)c * row_stride + q * 4;
#pragma unroll 4
    for (int s = 0; s < n_streams; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(row + (size_t)s * 2 * row_stride);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(bus + (size_t)c * bus_stride + q * 4) = acc;
}

}  // namespace ohs
