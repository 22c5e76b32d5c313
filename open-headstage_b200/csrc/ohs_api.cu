// ohs_api.cu — the C ABI of include/ohs.h over the kernels of ohs_kernels.cuh.
//
// Host-side state of one engine handle: the per-stream bindings (HRIR set, EQ set, gain), host copies of the
// impulse responses and EQ coefficients (uploaded lazily, on the next process call, so a burst of set_ir /
// update_band calls costs one upload — the reference recomputes all 20 biquads on every process() call,
// src/lib.rs:1180-1193), and the device-resident stream state: frequency-domain delay line, overlap-save block,
// biquad states.  No CPU fallback exists anywhere in this file: without a CUDA device ohs_create fails.
#include "../../include/ohs.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library itself is opened with dlopen by the first ohs_comm_* call

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ohs_aux_kernels.cuh"
#include "ohs_launch.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define OHS_CUDA(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess) return fail(OHS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                           __FILE__, __LINE__);                                                 \
    } while (0)

#define OHS_CHECK_HANDLE(h) \
    do { if (!(h)) return fail(OHS_ERR_INVALID, "null engine handle"); } while (0)

constexpr int kPipe = 3;  // staging buffers of the host-pointer path

}  // namespace

struct ohs_engine {
    ohs_config cfg{};
    int B = 0, N = 0, pmax = 1, G = 1;
    cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    bool timing = false, timed = false;   // ohs_enable_timing: event pair around the kernels of a process call

    // device
    int* d_stream_hrir = nullptr;
    int* d_stream_eq = nullptr;
    float* d_stream_gain = nullptr;
    float4* d_filt = nullptr;
    int* d_set_parts = nullptr;
    float2* d_fdl = nullptr;
    float2* d_prev = nullptr;
    float* d_eqc = nullptr;
    float4* d_eqs = nullptr;
    float2* d_tw = nullptr;
    float* d_ir = nullptr;
    int* d_set_list = nullptr;
    unsigned char* d_set_flags = nullptr;
    size_t filt_bytes = 0;

    // host mirrors
    std::vector<int> h_stream_hrir, h_stream_eq, h_set_parts, h_path_parts;
    std::vector<float> h_gain, h_eqc;
    std::vector<std::vector<float>> h_ir;  // [set*4 + path]
    std::vector<unsigned char> set_dirty, set_external;
    bool any_set_dirty = false, eqc_dirty = true, bind_dirty = true, gain_dirty = true;

    int head = 0;
    int eq_enable = 0, conv_enable = 1, bypass = 0;
    uint64_t launches = 0;
    // tuning switches, read from the environment ONCE at ohs_create (and settable through the API afterwards)
    int time_batch = 1;            // OHS_TIME_BATCH / ohs_set_time_batch: 0 keeps long responses on the block-by-block kernel
    size_t stage_bytes = (size_t)24 << 20;  // OHS_STAGE_MB: staging chunk of the host-pointer path
    bool dependent_launch = true;  // OHS_PDL=0 switches programmatic dependent launches off
    int latency_blocks = 2;        // OHS_LATENCY_BLOCKS: launches of up to this many blocks run the latency variant
    unsigned long long* d_trace = nullptr;  // ohs_debug_trace (OHS_TRACE builds)
    std::vector<float> h_ir_padded;         // [set][4][pmax*B] host mirror of d_ir, uploaded in one copy per commit

    // host-pointer path
    float* d_stage[kPipe] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[kPipe] = {}, ev_comp[kPipe] = {}, ev_out[kPipe] = {};
    size_t stage_frames = 0;

    // time-batched long-response path: spectra in time order and frequency-domain products of one sub-launch
    float2* d_zlin = nullptr;
    float2* d_wlin = nullptr;
    size_t zlin_blocks = 0;  // blocks per sub-launch the two buffers are sized for
    // The time-ordered buffer is circular (zl_cap = pmax-1 + zlin_blocks slots per stream; zl_base = slot of the next
    // block) and carries the convolution history from one time-batched call to the next; the delay-line ring d_fdl is
    // what the block-by-block kernel and the state blob use.  At least one of the two is current at any time; they are
    // converted into each other (gather / scatter, a copy of pmax-1 spectra per stream) only when the route changes.
    int zl_base = 0, zl_cap = 0;
    bool zlin_valid = false, ring_valid = true;
    // ... and its EQ pre-pass: the EQ-filtered rows of two sub-launches (halves used alternately), written by the EQ-only
    // kernel on its own stream while the previous chunk's transforms and products run on the engine's stream
    float* d_xf = nullptr;   // [2 halves][stream][2][zlin_blocks * B]
    cudaStream_t eq_stream = nullptr;
    std::vector<cudaEvent_t> ev_eq;          // EQ pre-pass of chunk c has finished (pool, used round-robin)
    cudaEvent_t ev_fwd[2] = {nullptr, nullptr};   // forward transforms have consumed a half of d_xf
    cudaEvent_t ev_call = nullptr;           // everything enqueued on the engine's stream before this call
    size_t eq_events_used = 0;
    int tb_overlap = 1;      // OHS_TB_OVERLAP=0: EQ pre-pass on the engine's stream, one chunk per sub-launch (A/B)
    int tb_eq_g = 0;         // OHS_TB_EQ_G: streams per CTA of the EQ pre-pass (A/B)
    int tb_eq_smem_kb = -1;  // OHS_TB_EQ_SMEM_KB: dynamic shared memory its CTAs ask for at least (A/B; -1: by shape)
    int tb_chunk = 0;        // OHS_TB_CHUNK=n: uniform chunks of n blocks instead of the ramped schedule (A/B)

    // FIFO adaptor (src/dsp/convolution.rs:141-182)
    std::vector<float> fifo_in, fifo_out;  // [row][cap]
    size_t fifo_cap = 0, fifo_in_len = 0, fifo_out_len = 0;
};

namespace {

using namespace ohs;

int launch_render(ohs_engine* h, const RenderParams& p, int first_stream = 0) {
    // few blocks per launch: nothing overlaps inside the launch, so the variant built for latency runs it
    const bool latency = p.n_blocks <= h->latency_blocks;
    RenderLaunch L{h->G, h->cfg.device, h->stream, first_stream, latency ? 1 : 0, 0, h->dependent_launch};
    RenderParams q = p;
    q.trace = h->d_trace;
    cudaError_t e = cudaErrorInvalidValue;
    switch (h->N) {
        case 128: e = render_launch_128(L, q); break;
        case 256: e = render_launch_256(L, q); break;
        case 512: e = render_launch_512(L, q); break;
        case 1024: e = render_launch_1024(L, q); break;
        case 2048: e = render_launch_2048(L, q); break;
        default: return fail(OHS_ERR_INVALID, "unsupported block size %d", h->B);
    }
    if (e == cudaErrorInvalidConfiguration) { cudaGetLastError(); return fail(OHS_ERR_INVALID, "%d streams per CTA do not fit for block %d", h->G, h->B); }
    if (e != cudaSuccess) return fail(OHS_ERR_CUDA, "render kernel launch failed: %s", cudaGetErrorString(e));
    OHS_CUDA(cudaGetLastError());
    h->launches++;
    return OHS_OK;
}

bool render_fits(int N, int G) {
    switch (N) {
        case 128: return render_fits_128(G); case 256: return render_fits_256(G); case 512: return render_fits_512(G);
        case 1024: return render_fits_1024(G); case 2048: return render_fits_2048(G);
    }
    return false;
}

// per-device "attribute already set" flags of the small kernels; handles on different threads may race here, hence
// atomic (setting an attribute twice is harmless)
struct AttrOnce {
    std::atomic<unsigned char> done[64];
    bool need(int dev) const { return dev >= 0 && dev < 64 && !done[dev].load(std::memory_order_acquire); }
    void mark(int dev) { done[dev].store(1, std::memory_order_release); }
};

template <int N> int launch_setup_n(ohs_engine* h, int max_parts, int n_sets) {
    using SM = SetupSmem<N>;
    static AttrOnce once;
    const int dev = h->cfg.device;
    if (once.need(dev)) {
        OHS_CUDA(cudaFuncSetAttribute(setup_filters_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::kBytes));
        once.mark(dev);
    }
    dim3 grid(max_parts, n_sets);
    setup_filters_kernel<N><<<grid, SM::T, SM::kBytes, h->stream>>>(h->d_ir, h->d_filt, h->d_tw, h->d_set_list, h->d_set_parts, h->pmax);
    OHS_CUDA(cudaGetLastError());
    h->launches++;
    return OHS_OK;
}

int launch_setup(ohs_engine* h, int max_parts, int n_sets) {
    switch (h->N) {
        case 128: return launch_setup_n<128>(h, max_parts, n_sets);
        case 256: return launch_setup_n<256>(h, max_parts, n_sets);
        case 512: return launch_setup_n<512>(h, max_parts, n_sets);
        case 1024: return launch_setup_n<1024>(h, max_parts, n_sets);
        case 2048: return launch_setup_n<2048>(h, max_parts, n_sets);
    }
    return fail(OHS_ERR_INVALID, "unsupported block size %d", h->B);
}

// the delay-line ring brought up to date from the time-ordered buffer (after time-batched calls), before anything reads
// or partially rewrites it
int ensure_ring(ohs_engine* h) {
    if (h->ring_valid) return OHS_OK;
    if (h->pmax > 1 && h->d_zlin) {
        scatter_history_kernel<<<dim3((unsigned)(h->pmax - 1), (unsigned)h->cfg.n_streams), 256, 0, h->stream>>>(
            h->d_zlin, h->d_fdl, h->N, h->pmax, h->head, (long long)h->zl_cap * h->N, h->zl_base, h->zl_cap);
        OHS_CUDA(cudaGetLastError());
        h->launches++;
    }
    h->ring_valid = true;
    return OHS_OK;
}

int clear_history(ohs_engine* h, bool only_flagged) {
    if (only_flagged) { const int rc = ensure_ring(h); if (rc) return rc; }
    h->ring_valid = true; h->zlin_valid = false;   // the cleared ring is the current history from here on
    clear_history_kernel<<<h->cfg.n_streams, 256, 0, h->stream>>>(h->d_fdl, h->d_prev, h->d_stream_hrir, h->cfg.n_streams,
                                                                 only_flagged ? h->d_set_flags : nullptr,
                                                                 (size_t)h->pmax * h->N, (size_t)h->B);
    OHS_CUDA(cudaGetLastError());
    h->launches++;
    return OHS_OK;
}

int upload_bindings(ohs_engine* h) {
    if (h->bind_dirty) {
        OHS_CUDA(cudaMemcpyAsync(h->d_stream_hrir, h->h_stream_hrir.data(), sizeof(int) * h->cfg.n_streams, cudaMemcpyHostToDevice, h->stream));
        OHS_CUDA(cudaMemcpyAsync(h->d_stream_eq, h->h_stream_eq.data(), sizeof(int) * h->cfg.n_streams, cudaMemcpyHostToDevice, h->stream));
        h->bind_dirty = false;
    }
    if (h->gain_dirty) {
        OHS_CUDA(cudaMemcpyAsync(h->d_stream_gain, h->h_gain.data(), sizeof(float) * h->cfg.n_streams, cudaMemcpyHostToDevice, h->stream));
        h->gain_dirty = false;
    }
    if (h->eqc_dirty) {
        OHS_CUDA(cudaMemcpyAsync(h->d_eqc, h->h_eqc.data(), sizeof(float) * h->h_eqc.size(), cudaMemcpyHostToDevice, h->stream));
        h->eqc_dirty = false;
    }
    // pageable host sources: the async copies above are staged by the runtime before returning
    return OHS_OK;
}

int commit_filters(ohs_engine* h) {
    int rc = upload_bindings(h);
    if (rc) return rc;
    if (!h->any_set_dirty) return OHS_OK;
    // every dirty set's taps go into the host mirror, ONE copy uploads the span of dirty sets and ONE launch transforms
    // them (config 4 binds a set per stream: thousands of sets per commit)
    const int n_sets = h->cfg.n_hrir_sets;
    const size_t per_path = (size_t)h->pmax * h->B, per_set = 4 * per_path;
    if (h->h_ir_padded.size() != per_set * n_sets) h->h_ir_padded.assign(per_set * n_sets, 0.f);
    std::vector<int> list;
    int max_parts = 1, lo = n_sets, hi = -1;
    for (int s = 0; s < n_sets; ++s) {
        if (!h->set_dirty[s]) continue;
        float* dst = h->h_ir_padded.data() + (size_t)s * per_set;
        std::fill(dst, dst + per_set, 0.f);
        int parts = 1;
        for (int p = 0; p < 4; ++p) {
            const std::vector<float>& ir = h->h_ir[(size_t)s * 4 + p];
            std::copy(ir.begin(), ir.end(), dst + p * per_path);
            parts = std::max(parts, h->h_path_parts[(size_t)s * 4 + p]);
        }
        h->h_set_parts[s] = parts;
        max_parts = std::max(max_parts, parts);
        lo = std::min(lo, s); hi = std::max(hi, s);
        list.push_back(s);
    }
    OHS_CUDA(cudaMemcpyAsync(h->d_ir + (size_t)lo * per_set, h->h_ir_padded.data() + (size_t)lo * per_set,
                             sizeof(float) * per_set * (size_t)(hi - lo + 1), cudaMemcpyHostToDevice, h->stream));
    OHS_CUDA(cudaMemcpyAsync(h->d_set_parts, h->h_set_parts.data(), sizeof(int) * n_sets, cudaMemcpyHostToDevice, h->stream));
    OHS_CUDA(cudaMemcpyAsync(h->d_set_list, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, h->stream));
    OHS_CUDA(cudaMemcpyAsync(h->d_set_flags, h->set_dirty.data(), n_sets, cudaMemcpyHostToDevice, h->stream));
    OHS_CUDA(cudaStreamSynchronize(h->stream));  // pageable sources (`list` is a local) must be consumed before returning
    rc = launch_setup(h, max_parts, (int)list.size());
    if (rc) return rc;
    // set_ir clears the history of the streams that use the set (src/dsp/convolution.rs:135-138)
    rc = clear_history(h, true);
    if (rc) return rc;
    std::fill(h->set_dirty.begin(), h->set_dirty.end(), 0);
    h->any_set_dirty = false;
    return OHS_OK;
}

// biquad 0.4.2 Coefficients::<f32>::from_params, f32 arithmetic and libm like the crate (src/dsp/parametric_eq.rs:105-111)
int eq_design_impl(int type, float fs, float fc, float q, float gain_db, float out[5]) {
    if (2.0f * fc > fs) return fail(OHS_ERR_NYQUIST, "EQ design: 2*fc (%g) > fs (%g)", 2.0 * fc, (double)fs);
    if (q < 0.0f) return fail(OHS_ERR_NEGATIVE_Q, "EQ design: negative Q %g", (double)q);
    const float pi = 3.14159265358979323846f;
    const float omega = 2.0f * pi * fc / fs;
    const float sn = sinf(omega), cs = cosf(omega);
    const float alpha = sn / (2.0f * q);
    float b0, b1, b2, a0, a1, a2;
    bool divide = false;
    if (type == OHS_FILTER_PEAK || type == OHS_FILTER_LOWSHELF || type == OHS_FILTER_HIGHSHELF) {
        const float a = powf(10.0f, gain_db / 40.0f);
        divide = true;
        if (type == OHS_FILTER_PEAK) {
            b0 = 1.0f + alpha * a; b1 = -2.0f * cs; b2 = 1.0f - alpha * a;
            a0 = 1.0f + alpha / a; a1 = -2.0f * cs; a2 = 1.0f - alpha / a;
        } else {
            const float beta = 2.0f * alpha * sqrtf(a);
            const float ap = a + 1.0f, am = a - 1.0f;
            if (type == OHS_FILTER_LOWSHELF) {
                b0 = a * (ap - am * cs + beta); b1 = 2.0f * a * (am - ap * cs); b2 = a * (ap - am * cs - beta);
                a0 = ap + am * cs + beta; a1 = -2.0f * (am + ap * cs); a2 = ap + am * cs - beta;
            } else {
                b0 = a * (ap + am * cs + beta); b1 = -2.0f * a * (am + ap * cs); b2 = a * (ap + am * cs - beta);
                a0 = ap - am * cs + beta; a1 = 2.0f * (am - ap * cs); a2 = ap - am * cs - beta;
            }
        }
    } else {
        a0 = 1.0f + alpha; a1 = -2.0f * cs; a2 = 1.0f - alpha;
        switch (type) {
            case OHS_FILTER_LOWPASS: b0 = (1.0f - cs) * 0.5f; b1 = 1.0f - cs; b2 = (1.0f - cs) * 0.5f; break;
            case OHS_FILTER_HIGHPASS: b0 = (1.0f + cs) * 0.5f; b1 = -(1.0f + cs); b2 = (1.0f + cs) * 0.5f; break;
            case OHS_FILTER_BANDPASS: b0 = sn / 2.0f; b1 = 0.0f; b2 = -(sn / 2.0f); break;
            case OHS_FILTER_NOTCH: b0 = 1.0f; b1 = -2.0f * cs; b2 = 1.0f; break;
            case OHS_FILTER_ALLPASS: b0 = 1.0f - alpha; b1 = -2.0f * cs; b2 = 1.0f + alpha; break;
            default: return fail(OHS_ERR_INVALID, "unknown filter type %d", type);
        }
    }
    if (divide) {
        out[0] = b0 / a0; out[1] = b1 / a0; out[2] = b2 / a0; out[3] = a1 / a0; out[4] = a2 / a0;
    } else {
        const float inv = 1.0f / a0;
        out[0] = b0 * inv; out[1] = b1 * inv; out[2] = b2 * inv; out[3] = a1 * inv; out[4] = a2 * inv;
    }
    return OHS_OK;
}

int set_band_host(ohs_engine* h, int eq_set, int band, const float c[5], int enabled) {
    if (eq_set < 0 || eq_set >= h->cfg.n_eq_sets) return fail(OHS_ERR_INVALID, "eq_set %d out of range", eq_set);
    if (band < 0 || band >= h->cfg.n_bands) return OHS_OK;  // silently ignored, src/dsp/parametric_eq.rs:145
    float* dst = h->h_eqc.data() + ((size_t)eq_set * kMaxBands + band) * kEqCoefStride;
    for (int i = 0; i < 5; ++i) dst[i] = c[i];
    dst[5] = enabled ? 1.0f : 0.0f;
    h->eqc_dirty = true;
    return OHS_OK;
}

int check_audio_args(ohs_engine* h, const void* in, const void* out, size_t n_frames, size_t row_stride, bool device) {
    if (!in || !out) return fail(OHS_ERR_INVALID, "null audio pointer");
    if (h->conv_enable && n_frames % (size_t)h->B) return fail(OHS_ERR_INVALID, "n_frames %zu is not a multiple of the engine block %d (use ohs_process_fifo)", n_frames, h->B);
    if (row_stride < n_frames) return fail(OHS_ERR_INVALID, "row_stride %zu < n_frames %zu", row_stride, n_frames);
    if (device && (((uintptr_t)in | (uintptr_t)out) & 15u || (row_stride & 3u)))
        return fail(OHS_ERR_ALIGNMENT, "device audio pointers must be 16-byte aligned and row_stride a multiple of 4 frames");
    return OHS_OK;
}

// called once, from ohs_create (the only place the environment is consulted)
// frames per staging chunk of the host-pointer path: whole blocks, about stage_bytes per buffer, at least one block
size_t stage_chunk_frames(const ohs_engine* h, size_t n_frames) {
    const size_t rows = (size_t)h->cfg.n_streams * 2;
    size_t chunk = h->stage_bytes / (rows * sizeof(float));
    chunk = std::max<size_t>(h->B, (chunk / h->B) * h->B);
    return std::min(chunk, (n_frames + 3) / 4 * 4);
}

int ensure_stage_buffers(ohs_engine* h, size_t chunk) {
    if (chunk <= h->stage_frames) return OHS_OK;
    const size_t rows = (size_t)h->cfg.n_streams * 2;
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    OHS_CUDA(cudaStreamSynchronize(h->d2h));
    for (int i = 0; i < kPipe; ++i) {
        if (h->d_stage[i]) OHS_CUDA(cudaFree(h->d_stage[i]));
        h->d_stage[i] = nullptr;
        OHS_CUDA(cudaMalloc(&h->d_stage[i], rows * chunk * sizeof(float)));
    }
    h->stage_frames = chunk;
    return OHS_OK;
}

int pick_streams_per_cta(const ohs_engine* h) {
    if (const char* e = getenv("OHS_STREAMS_PER_CTA")) {
        const int g = atoi(e);
        if (g >= 1 && g <= kMaxG && render_fits(h->N, g)) return g;
    }
    cudaDeviceProp prop{};
    int sms = 148;
    if (cudaGetDeviceProperties(&prop, h->cfg.device) == cudaSuccess) sms = prop.multiProcessorCount;
    // Few streams per SM (config 2: 1024 streams -> 7 per CTA, 147 CTAs): one CTA per SM holding ceil(n_streams / SMs)
    // streams, so every SM gets the same share.  Many streams per SM (config 3: 55 per SM): CTAs of three streams, whose
    // six chains exactly fill one EQ warp, several CTAs resident per SM (measured best on config 3).
    int g = (h->cfg.n_streams + sms - 1) / sms;
    g = (g > kMaxG) ? 3 : std::max(1, g);
    while (g > 1 && !render_fits(h->N, g)) --g;
    return g;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kTimeBatch = 8;  // consecutive blocks a bin_conv_kernel thread accumulates (TB)

template <int N> int launch_inverse(ohs_engine* h, const float2* d_w, float* d_out, int K, size_t row_stride) {
    constexpr size_t smem = sizeof(float2) * 2 * padded_len(N);
    static AttrOnce once;
    const int dev = h->cfg.device;
    if (smem > 48 * 1024 && once.need(dev)) {
        OHS_CUDA(cudaFuncSetAttribute(inverse_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        once.mark(dev);
    }
    inverse_kernel<N><<<dim3(K, h->cfg.n_streams), SetupSmem<N>::TX, smem, h->stream>>>(d_w, d_out, h->d_tw, h->d_stream_gain, K,
                                                                                      (long long)row_stride);
    OHS_CUDA(cudaGetLastError());
    h->launches++;
    return OHS_OK;
}

// forward transforms of blocks [t0, t0 + k) of a sub-launch from the filtered rows xf (block 0's history: d_prev)
template <int N> int launch_forward(ohs_engine* h, const float* xf, long long xf_stride, int t0, int k, long long zstride, int zbase) {
    // zbase: the time-ordered buffer's slot of the sub-launch's block 0
    constexpr size_t smem = sizeof(float2) * 2 * padded_len(N);
    static AttrOnce once;
    const int dev = h->cfg.device;
    if (smem > 48 * 1024 && once.need(dev)) {
        OHS_CUDA(cudaFuncSetAttribute(forward_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        once.mark(dev);
    }
    forward_kernel<N><<<dim3(k, h->cfg.n_streams), SetupSmem<N>::TX, smem, h->stream>>>(
        xf, xf_stride, t0, reinterpret_cast<const float*>(h->d_prev), h->d_zlin, zstride, zbase, h->zl_cap, h->d_tw);
    OHS_CUDA(cudaGetLastError());
    h->launches++;
    return OHS_OK;
}

// Long responses, many blocks per call.  Per sub-launch of up to `zlin_blocks` blocks:
//   (0) EQ pre-pass: the EQ-only variant of the render kernel (a thin CTA per three streams: two EQ warps, 38 KB) filters
//       the input rows into d_xf on its OWN stream.  The biquad chain is strictly sequential — 1024 steps per block at
//       the chain's latency, with most of the SM idle — while everything below is throughput-bound and independent of
//       later input, so the pre-pass of chunk c+1 runs beside the transforms and products of chunk c;
//   (1) delay-line history into the time-ordered buffer (once per sub-launch);
//   and per chunk (8, 8, 16, 32 blocks in a call's first sub-launch, so that only the first 8 blocks' EQ is exposed; the
//   whole sub-launch afterwards, when the pre-pass is running ahead anyway):
//   (2) forward transforms (spectra to the ring and to the time-ordered buffer), (3) the per-bin convolution along time,
//   (4) the inverse transforms.
// The delay-line ring, the overlap-save block and the EQ state end up exactly where the block-by-block path leaves
// them, so the two can be mixed freely between calls.
// whether a call of n_blocks takes the time-batched route
bool time_batch_eligible(const ohs_engine* h, int n_blocks) {
    return h->time_batch && h->conv_enable && h->pmax >= 8 && n_blocks >= kTimeBatch && h->cfg.n_streams <= 65535;  // (grid.y = stream)
}

// Scratch of the time-batched route for calls of n_blocks (ohs_prepare calls this ahead of time so that no allocation
// happens inside a timed call).  1: does not fit the 2 GiB budget (the caller stays on the block-by-block kernel).
int ensure_time_batch_scratch(ohs_engine* h, int n_blocks) {
    const size_t S = (size_t)h->cfg.n_streams, N = (size_t)h->N, hist = (size_t)h->pmax - 1;
    const size_t budget = (size_t)2 << 30;
    // per block: a spectrum and a product (8N bytes each) and the filtered rows in both halves of d_xf (2 * 2 * 4B = 8N)
    const size_t per_block = 3 * S * N * sizeof(float2), fixed = S * hist * N * sizeof(float2);
    size_t kc = fixed < budget ? (budget - fixed) / per_block : 0;
    kc = std::min<size_t>(std::min<size_t>(kc, 64), (size_t)n_blocks);
    if (kc < (size_t)kTimeBatch) return 1;
    if (kc > h->zlin_blocks) {
        { const int rc = ensure_ring(h); if (rc) return rc; }   // the history moves to the ring before its buffer goes
        h->zlin_valid = false;
        OHS_CUDA(cudaStreamSynchronize(h->stream));
        if (h->eq_stream) OHS_CUDA(cudaStreamSynchronize(h->eq_stream));
        if (h->d_zlin) OHS_CUDA(cudaFree(h->d_zlin));
        if (h->d_wlin) OHS_CUDA(cudaFree(h->d_wlin));
        if (h->d_xf) OHS_CUDA(cudaFree(h->d_xf));
        h->d_zlin = h->d_wlin = nullptr; h->d_xf = nullptr; h->zlin_blocks = 0;
        OHS_CUDA(cudaMalloc(&h->d_zlin, S * (hist + kc) * N * sizeof(float2)));
        OHS_CUDA(cudaMalloc(&h->d_wlin, S * kc * N * sizeof(float2)));
        OHS_CUDA(cudaMalloc(&h->d_xf, 2 * S * 2 * kc * (size_t)h->B * sizeof(float)));
        h->zlin_blocks = kc;
        h->zl_cap = (int)(hist + kc); h->zl_base = (int)hist;
    }
    if (!h->eq_stream) {
        int lo = 0, hi = 0;
        OHS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        OHS_CUDA(cudaStreamCreateWithPriority(&h->eq_stream, cudaStreamNonBlocking, hi));   // its thin CTAs go first
        OHS_CUDA(cudaEventCreateWithFlags(&h->ev_fwd[0], cudaEventDisableTiming));
        OHS_CUDA(cudaEventCreateWithFlags(&h->ev_fwd[1], cudaEventDisableTiming));
        OHS_CUDA(cudaEventCreateWithFlags(&h->ev_call, cudaEventDisableTiming));
    }
    return OHS_OK;
}

// the per-bin convolution of the k blocks from slot zslot0 of the circular time-ordered buffer on, products to `w`
// ([stream][k][N])
int launch_bin_conv(ohs_engine* h, int zslot0, float2* w, int k, long long zstride) {
    const size_t S = (size_t)h->cfg.n_streams, N = (size_t)h->N;
    const dim3 blk(128);
    const unsigned gx = (unsigned)((N / 2 + 31) / 32), gz = (unsigned)((S + 3) / 4);
    static AttrOnce once;
    const int dev = h->cfg.device;
    if (once.need(dev)) {
        OHS_CUDA(cudaFuncSetAttribute(bin_conv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 16 * 128 * sizeof(float4))));
        once.mark(dev);
    }
    // 16 blocks per thread (half the spectrum and filter traffic per FMA) from 16 blocks per launch on: measured faster
    // than 8 at K = 16 (0.529 vs 0.560 ms), 32, 64 and 128 on config 5
    if (k >= 16)
        bin_conv_kernel<16><<<dim3(gx, (unsigned)((k + 15) / 16), gz), blk, 2 * 16 * 128 * sizeof(float4), h->stream>>>(
            h->d_zlin, w, h->d_filt, h->d_stream_hrir, h->d_set_parts, (int)N, h->pmax, k, (int)S, zstride, zslot0, h->zl_cap);
    else
        bin_conv_kernel<kTimeBatch><<<dim3(gx, (unsigned)((k + kTimeBatch - 1) / kTimeBatch), gz), blk, 2 * kTimeBatch * 128 * sizeof(float4), h->stream>>>(
            h->d_zlin, w, h->d_filt, h->d_stream_hrir, h->d_set_parts, (int)N, h->pmax, k, (int)S, zstride, zslot0, h->zl_cap);
    OHS_CUDA(cudaGetLastError());
    h->launches++;
    return OHS_OK;
}

// EQ pre-pass of `frames` frames of every stream on `stream`: rows in (stride in_stride) -> rows out (stride out_stride)
int launch_eq_prepass(ohs_engine* h, const RenderParams& base, const float* d_in, long long in_stride, float* d_out, long long out_stride,
                      size_t frames, cudaStream_t stream) {
    constexpr int kB = 256;   // the EQ-only variant exists for N = 512; the chain does not care how its rows are cut into blocks
    RenderParams q = base;
    q.in = d_in; q.out = d_out; q.row_stride = in_stride; q.out_row_stride = out_stride;
    q.n_blocks = (int)((frames + kB - 1) / kB);
    q.tail_frames = (int)(frames - (size_t)(q.n_blocks - 1) * kB);
    q.conv_enable = 0; q.eq_enable = 1; q.filt_in_smem = 0;
    q.trace = nullptr;
    // Shape of the pre-pass (RenderSmem, V = 2).  Many streams, run beside the previous chunk's kernels: six per CTA — four
    // EQ warps, one per scheduler partition — and 200 KB of shared memory asked for, so that each CTA OWNS its SM: next to
    // the per-bin kernel's FMA-saturated warps every instruction of the biquad chain waits for its issue slot and the
    // chain, the route's critical path, runs 2.5x slower (measured; config 5 at 256 blocks per call: 4.95 ms with 43 SMs
    // owned, 5.65 ms with thin CTAs of three streams sharing 86 SMs, 5.82 ms with twelve streams = two EQ warps per
    // partition on 22 SMs).  Few streams or one stream of work: thin CTAs of three.
    const bool own_sm = stream != h->stream && h->cfg.n_streams >= 12;
    const int g = (h->tb_eq_g > 0) ? h->tb_eq_g : (own_sm ? 6 : 3);
    const size_t min_smem = (h->tb_eq_smem_kb >= 0) ? (size_t)h->tb_eq_smem_kb << 10 : (own_sm ? (size_t)200 << 10 : 0);
    // no programmatic dependent launch here: the next chunk's CTAs, started early, would sit on SMs of their own (they
    // cannot share one with this chunk's) and take them from the transforms of the chunk before
    RenderLaunch L{g, h->cfg.device, stream, 0, 2, min_smem, false};
    const cudaError_t e = render_launch_512(L, q);
    if (e != cudaSuccess) return fail(OHS_ERR_CUDA, "EQ pre-pass launch failed: %s", cudaGetErrorString(e));
    h->launches++;
    return OHS_OK;
}

int process_time_batched(ohs_engine* h, RenderParams p, const float* d_in, float* d_out, size_t row_stride) {
    const size_t S = (size_t)h->cfg.n_streams, N = (size_t)h->N, hist = (size_t)h->pmax - 1, B = (size_t)h->B;
    {
        const int rc = ensure_time_batch_scratch(h, p.n_blocks);
        if (rc) return rc;   // 1: caller falls back to the block-by-block kernel
    }
    const size_t kc = h->zlin_blocks;
    const long long zstride = (long long)((hist + kc) * N);
    const long long xstride = (long long)(kc * B);               // frames between the rows of a half of d_xf
    const size_t xhalf = S * 2 * kc * B;                         // floats per half
    const int total = p.n_blocks;
    const bool eq_on = p.eq_enable != 0;
    const bool overlap = eq_on && h->tb_overlap;
    cudaStream_t eqs = overlap ? h->eq_stream : h->stream;
    const int n_sub = (total + (int)kc - 1) / (int)kc;
    // chunks of sub-launch j: (first block inside the sub-launch, blocks).  The EQ pre-pass of a chunk runs beside the
    // transforms of the chunk before it, so what a call cannot hide is the pre-pass of its first chunk and the transforms of
    // its last: the first sub-launch ramps up (16, 16, 32), the last one ramps down (32, 16, 16), everything between goes
    // in the largest pieces.  (Chunks of 8 were measured and lose: the per-bin kernel's 8-block form takes 128 us per
    // chunk against 187 us for 16 blocks; config 5 at 64 blocks per call: 1.39 ms with 16-16-16-16, 1.50 ms with
    // 8-8-16-16-8-8, 1.61 ms un-overlapped.)
    auto chunks_of = [&](int j, std::vector<std::pair<int, int>>& out) {
        out.clear();
        const int k = std::min<int>((int)kc, total - j * (int)kc);
        if (!overlap) { out.emplace_back(0, k); return; }
        if (h->tb_chunk > 0) {
            for (int at = 0; at < k; at += h->tb_chunk) out.emplace_back(at, std::min(h->tb_chunk, k - at));
            return;
        }
        std::vector<int> head, tail;
        int rest = k;
        if (j == 0) for (int c : {16, 16, 32}) if (rest >= 2 * c) { head.push_back(c); rest -= c; }
        if (j == n_sub - 1) for (int c : {16, 16, 32}) if (rest >= 2 * c) { tail.push_back(c); rest -= c; }
        int at = 0;
        for (int c : head) { out.emplace_back(at, c); at += c; }
        if (rest > 0) { out.emplace_back(at, rest); at += rest; }
        for (size_t i = tail.size(); i-- > 0;) { out.emplace_back(at, tail[i]); at += tail[i]; }
    };
#ifdef OHS_TB_TRACE   // debug builds only: a timing event behind every kernel of the route, printed as a timeline per call
    struct Mark { cudaEvent_t ev; const char* what; int a, b; };
    std::vector<Mark> marks;
    auto mark = [&](cudaStream_t st, const char* what, int a, int b) {
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); marks.push_back({e, what, a, b});
    };
#define OHS_TB_MARK(st, what, a, b) mark(st, what, a, b)
#else
#define OHS_TB_MARK(st, what, a, b) do { } while (0)
#endif
    OHS_TB_MARK(h->stream, "call", 0, 0);
    std::vector<std::vector<std::pair<int, int>>> chunks(n_sub);
    std::vector<std::vector<cudaEvent_t>> done_ev(n_sub);
    for (int j = 0; j < n_sub; ++j) chunks_of(j, chunks[j]);
    auto next_event = [&](cudaEvent_t* ev) -> int {
        if (h->eq_events_used == h->ev_eq.size()) {
            cudaEvent_t e;
            OHS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->ev_eq.push_back(e);
        }
        *ev = h->ev_eq[h->eq_events_used++];
        return OHS_OK;
    };
    h->eq_events_used = 0;
    // enqueue the EQ pre-pass of sub-launch j (its chunks back to back on the EQ stream, an event behind each)
    auto enqueue_eq = [&](int j) -> int {
        if (!eq_on) return OHS_OK;
        float* xf = h->d_xf + (size_t)(j & 1) * xhalf;
        if (overlap && j >= 2) OHS_CUDA(cudaStreamWaitEvent(eqs, h->ev_fwd[j & 1], 0));   // sub-launch j-2's transforms have read this half
        for (const auto& c : chunks[j]) {
            const size_t f0 = ((size_t)j * kc + c.first) * B;
            int rc = launch_eq_prepass(h, p, d_in + f0, (long long)row_stride, xf + (size_t)c.first * B, xstride, (size_t)c.second * B, eqs);
            if (rc) return rc;
            OHS_TB_MARK(eqs, "eq", j, c.first);
            if (overlap) {
                cudaEvent_t ev;
                rc = next_event(&ev);
                if (rc) return rc;
                OHS_CUDA(cudaEventRecord(ev, eqs));
                done_ev[j].push_back(ev);
            }
        }
        return OHS_OK;
    };
    if (overlap) {
        // the EQ stream starts behind everything already enqueued on the engine's stream (EQ state, input rows)
        OHS_CUDA(cudaEventRecord(h->ev_call, h->stream));
        OHS_CUDA(cudaStreamWaitEvent(eqs, h->ev_call, 0));
        for (int j = 0; j < std::min(2, n_sub); ++j) { const int rc = enqueue_eq(j); if (rc) return rc; }
    }
    for (int j = 0; j < n_sub; ++j) {
        const int k = std::min<int>((int)kc, total - j * (int)kc);
        if (!overlap) { const int rc = enqueue_eq(j); if (rc) return rc; }
        if (!h->zlin_valid) {
            // the route changes here (first call, or block-by-block calls since the last time-batched one): the ring's
            // history goes behind the next block's slot; between time-batched sub-launches and calls nothing is copied
            gather_history_kernel<<<dim3((unsigned)hist, (unsigned)S), 256, 0, h->stream>>>(h->d_fdl, h->d_zlin, (int)N, h->pmax, h->head, zstride,
                                                                                        h->zl_base, h->zl_cap);
            OHS_CUDA(cudaGetLastError());
            h->launches++;
            h->zlin_valid = true;
            OHS_TB_MARK(h->stream, "gather", j, 0);
        }
        const int zb = h->zl_base;   // slot of this sub-launch's block 0
        // filtered rows of this sub-launch: a half of d_xf (with the EQ off, a copy of the input rows: the transforms read
        // a block's predecessor too, which an in-place call has overwritten with output by then)
        const float* xf = h->d_xf + (size_t)(j & 1) * xhalf;
        const long long xs = xstride;
        if (!eq_on)
            OHS_CUDA(cudaMemcpy2DAsync(h->d_xf + (size_t)(j & 1) * xhalf, (size_t)xstride * sizeof(float), d_in + (size_t)j * kc * B,
                                       row_stride * sizeof(float), (size_t)k * B * sizeof(float), S * 2, cudaMemcpyDeviceToDevice, h->stream));
        float* out_j = d_out + (size_t)j * kc * B;
        for (size_t ci = 0; ci < chunks[j].size(); ++ci) {
            const int c0 = chunks[j][ci].first, ck = chunks[j][ci].second;
            if (overlap) OHS_CUDA(cudaStreamWaitEvent(h->stream, done_ev[j][ci], 0));
            int rc;
            switch (h->N) {
                case 128: rc = launch_forward<128>(h, xf, xs, c0, ck, zstride, zb); break;
                case 256: rc = launch_forward<256>(h, xf, xs, c0, ck, zstride, zb); break;
                case 512: rc = launch_forward<512>(h, xf, xs, c0, ck, zstride, zb); break;
                case 1024: rc = launch_forward<1024>(h, xf, xs, c0, ck, zstride, zb); break;
                case 2048: rc = launch_forward<2048>(h, xf, xs, c0, ck, zstride, zb); break;
                default: rc = fail(OHS_ERR_INVALID, "unsupported transform size %d", h->N);
            }
            if (rc) return rc;
            OHS_TB_MARK(h->stream, "forward", j, c0);
            float2* w = h->d_wlin + S * (size_t)c0 * N;
            rc = launch_bin_conv(h, (zb + c0) % h->zl_cap, w, ck, zstride);
            if (rc) return rc;
            OHS_TB_MARK(h->stream, "bin_conv", j, c0);
            float* out_c = out_j + (size_t)c0 * B;
            switch (h->N) {
                case 128: rc = launch_inverse<128>(h, w, out_c, ck, row_stride); break;
                case 256: rc = launch_inverse<256>(h, w, out_c, ck, row_stride); break;
                case 512: rc = launch_inverse<512>(h, w, out_c, ck, row_stride); break;
                case 1024: rc = launch_inverse<1024>(h, w, out_c, ck, row_stride); break;
                case 2048: rc = launch_inverse<2048>(h, w, out_c, ck, row_stride); break;
                default: rc = fail(OHS_ERR_INVALID, "unsupported transform size %d", h->N);
            }
            if (rc) return rc;
            OHS_TB_MARK(h->stream, "inverse", j, c0);
        }
        // overlap-save history of the next sub-launch or call: this one's last filtered block
        OHS_CUDA(cudaMemcpy2DAsync(h->d_prev, B * sizeof(float), xf + (size_t)(k - 1) * B, (size_t)xs * sizeof(float), B * sizeof(float),
                                   S * 2, cudaMemcpyDeviceToDevice, h->stream));
        if (overlap) {
            OHS_CUDA(cudaEventRecord(h->ev_fwd[j & 1], h->stream));
            if (j + 2 < n_sub) { const int rc = enqueue_eq(j + 2); if (rc) return rc; }
        }
        h->head = (int)(((size_t)h->head + k) % (size_t)h->pmax);
        h->zl_base = (zb + k) % h->zl_cap;
        h->ring_valid = false;   // the spectra of these blocks exist in the time-ordered buffer only (ensure_ring)
    }
#ifdef OHS_TB_TRACE
    cudaStreamSynchronize(h->stream);
    if (h->eq_stream) cudaStreamSynchronize(h->eq_stream);
    for (const Mark& m : marks) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, marks[0].ev, m.ev);
        fprintf(stderr, "tb_trace %-8s sub %d block %2d done at %8.1f us\n", m.what, m.a, m.b, ms * 1e3);
    }
    for (const Mark& m : marks) cudaEventDestroy(m.ev);
#endif
    return OHS_OK;
}

}  // namespace

extern "C" {

int ohs_abi_version(void) { return OHS_ABI_VERSION; }
const char* ohs_last_error(void) { return g_last_error.c_str(); }

int ohs_eq_design(int filter_type, float fs, float fc, float q, float gain_db, float out[5]) {
    if (!out) return fail(OHS_ERR_INVALID, "null output");
    return eq_design_impl(filter_type, fs, fc, q, gain_db, out);
}

int ohs_create(const ohs_config* cfg, ohs_engine** out) {
    if (!cfg || !out) return fail(OHS_ERR_INVALID, "null argument");
    *out = nullptr;
    const int B = cfg->block;
    if (!(B == 64 || B == 128 || B == 256 || B == 512 || B == 1024)) return fail(OHS_ERR_INVALID, "block must be 64, 128, 256, 512 or 1024 (got %d)", B);
    if (cfg->n_streams < 1) return fail(OHS_ERR_INVALID, "n_streams must be >= 1");
    if (cfg->n_bands < 0 || cfg->n_bands > OHS_MAX_BANDS) return fail(OHS_ERR_INVALID, "n_bands must be 0..%d", OHS_MAX_BANDS);
    if (cfg->n_hrir_sets < 1 || cfg->n_eq_sets < 1) return fail(OHS_ERR_INVALID, "need at least one HRIR set and one EQ set");
    if (cfg->max_taps < 1) return fail(OHS_ERR_INVALID, "max_taps must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(OHS_ERR_NO_DEVICE, "no CUDA device available; this engine has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(OHS_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    OHS_CUDA(cudaSetDevice(cfg->device));

    ohs_engine* h = new ohs_engine();
    h->cfg = *cfg;
    h->B = B;
    h->N = 2 * B;
    h->pmax = (cfg->max_taps + B - 1) / B;
    h->G = pick_streams_per_cta(h);
    if (const char* e = getenv("OHS_TIME_BATCH")) h->time_batch = atoi(e) != 0;
    if (const char* e = getenv("OHS_TB_OVERLAP")) h->tb_overlap = atoi(e) != 0;
    if (const char* e = getenv("OHS_TB_EQ_G")) { const int g = atoi(e); if (g == 3 || g == 6) h->tb_eq_g = g; }
    if (const char* e = getenv("OHS_TB_EQ_SMEM_KB")) { const int k = atoi(e); if (k >= 0 && k <= 227) h->tb_eq_smem_kb = k; }
    if (const char* e = getenv("OHS_TB_CHUNK")) { const int c = atoi(e); if (c >= 1 && c <= 64) h->tb_chunk = c; }
    if (const char* e = getenv("OHS_STAGE_MB")) { const long mb = atol(e); if (mb >= 1 && mb <= 4096) h->stage_bytes = (size_t)mb << 20; }
    if (const char* e = getenv("OHS_PDL")) h->dependent_launch = atoi(e) != 0;
    if (const char* e = getenv("OHS_LATENCY_BLOCKS")) h->latency_blocks = atoi(e);
    const int S = cfg->n_streams;
    const size_t per_path = (size_t)h->pmax * B;

    h->h_stream_hrir.assign(S, 0);
    h->h_stream_eq.assign(S, 0);
    h->h_gain.assign(S, 1.0f);
    h->h_set_parts.assign(cfg->n_hrir_sets, 1);
    h->h_path_parts.assign((size_t)cfg->n_hrir_sets * 4, 1);
    h->h_ir.assign((size_t)cfg->n_hrir_sets * 4, std::vector<float>());
    h->set_dirty.assign(cfg->n_hrir_sets, 0);
    h->set_external.assign(cfg->n_hrir_sets, 0);
    h->h_eqc.assign((size_t)cfg->n_eq_sets * kMaxBands * kEqCoefStride, 0.f);
    {
        // BiquadFilter::new: PeakingEQ(0 dB) @ 20 Hz, Q 0.707, disabled (src/dsp/parametric_eq.rs:63-76)
        float c[5] = {1.f, 0.f, 0.f, 0.f, 0.f};
        if (2.0f * 20.0f <= cfg->sample_rate) eq_design_impl(OHS_FILTER_PEAK, cfg->sample_rate, 20.0f, 0.707f, 0.0f, c);
        for (int e = 0; e < cfg->n_eq_sets; ++e)
            for (int b = 0; b < kMaxBands; ++b) {
                float* dst = h->h_eqc.data() + ((size_t)e * kMaxBands + b) * kEqCoefStride;
                for (int i = 0; i < 5; ++i) dst[i] = c[i];
            }
    }

#define OHS_TRY(expr)                                   \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) {                        \
            fail(OHS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
            ohs_destroy(h);                             \
            return OHS_ERR_CUDA;                        \
        }                                               \
    } while (0)
    OHS_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    OHS_TRY(cudaStreamCreateWithFlags(&h->h2d, cudaStreamNonBlocking));
    OHS_TRY(cudaStreamCreateWithFlags(&h->d2h, cudaStreamNonBlocking));
    OHS_TRY(cudaEventCreate(&h->ev_k0));
    OHS_TRY(cudaEventCreate(&h->ev_k1));
    for (int i = 0; i < kPipe; ++i) {
        OHS_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        OHS_TRY(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
        OHS_TRY(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }
    h->filt_bytes = sizeof(float4) * (size_t)cfg->n_hrir_sets * h->pmax * h->N;
    OHS_TRY(cudaMalloc(&h->d_stream_hrir, sizeof(int) * S));
    OHS_TRY(cudaMalloc(&h->d_stream_eq, sizeof(int) * S));
    OHS_TRY(cudaMalloc(&h->d_stream_gain, sizeof(float) * S));
    OHS_TRY(cudaMalloc(&h->d_filt, h->filt_bytes));
    OHS_TRY(cudaMalloc(&h->d_set_parts, sizeof(int) * cfg->n_hrir_sets));
    OHS_TRY(cudaMalloc(&h->d_set_list, sizeof(int) * cfg->n_hrir_sets));
    OHS_TRY(cudaMalloc(&h->d_set_flags, cfg->n_hrir_sets));
    OHS_TRY(cudaMalloc(&h->d_fdl, sizeof(float2) * (size_t)S * h->pmax * h->N));
    OHS_TRY(cudaMalloc(&h->d_prev, sizeof(float2) * (size_t)S * B));
    OHS_TRY(cudaMalloc(&h->d_eqc, sizeof(float) * h->h_eqc.size()));
    OHS_TRY(cudaMalloc(&h->d_eqs, sizeof(float4) * (size_t)S * kMaxBands));
    OHS_TRY(cudaMalloc(&h->d_tw, sizeof(float2) * h->N));
    OHS_TRY(cudaMalloc(&h->d_ir, sizeof(float) * (size_t)cfg->n_hrir_sets * 4 * per_path));
    OHS_TRY(cudaMemsetAsync(h->d_filt, 0, h->filt_bytes, h->stream));  // default IR = silence (src/dsp/convolution.rs:44-65)
    OHS_TRY(cudaMemsetAsync(h->d_fdl, 0, sizeof(float2) * (size_t)S * h->pmax * h->N, h->stream));
    OHS_TRY(cudaMemsetAsync(h->d_prev, 0, sizeof(float2) * (size_t)S * B, h->stream));
    OHS_TRY(cudaMemsetAsync(h->d_eqs, 0, sizeof(float4) * (size_t)S * kMaxBands, h->stream));
    OHS_TRY(cudaMemsetAsync(h->d_ir, 0, sizeof(float) * (size_t)cfg->n_hrir_sets * 4 * per_path, h->stream));
    {
        std::vector<float2> tw(h->N);
        switch (h->N) {
            case 128: fill_twiddles<128>(tw.data()); break;
            case 256: fill_twiddles<256>(tw.data()); break;
            case 512: fill_twiddles<512>(tw.data()); break;
            case 1024: fill_twiddles<1024>(tw.data()); break;
            case 2048: fill_twiddles<2048>(tw.data()); break;
        }
        OHS_TRY(cudaMemcpyAsync(h->d_tw, tw.data(), sizeof(float2) * h->N, cudaMemcpyHostToDevice, h->stream));
        OHS_TRY(cudaMemcpyAsync(h->d_set_parts, h->h_set_parts.data(), sizeof(int) * cfg->n_hrir_sets, cudaMemcpyHostToDevice, h->stream));
        OHS_TRY(cudaStreamSynchronize(h->stream));
    }
#undef OHS_TRY
    *out = h;
    return OHS_OK;
}

int ohs_destroy(ohs_engine* h) {
    if (!h) return OHS_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->eq_stream) { cudaStreamSynchronize(h->eq_stream); cudaStreamDestroy(h->eq_stream); }
    for (cudaEvent_t e : h->ev_eq) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_fwd[0], h->ev_fwd[1], h->ev_call}) if (e) cudaEventDestroy(e);
    void* ptrs[] = {h->d_stream_hrir, h->d_stream_eq, h->d_stream_gain, h->d_filt, h->d_set_parts, h->d_set_list, h->d_set_flags,
                    h->d_fdl, h->d_prev, h->d_eqc, h->d_eqs, h->d_tw, h->d_ir, h->d_stage[0], h->d_stage[1], h->d_stage[2],
                    h->d_zlin, h->d_wlin, h->d_xf};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < kPipe; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->ev_k0) cudaEventDestroy(h->ev_k0);
    if (h->ev_k1) cudaEventDestroy(h->ev_k1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->h2d) cudaStreamDestroy(h->h2d);
    if (h->d2h) cudaStreamDestroy(h->d2h);
    delete h;
    return OHS_OK;
}

// ---- HRIR ---------------------------------------------------------------------------------------------------
int ohs_set_ir(ohs_engine* h, int hrir_set, int path, const float* ir, size_t len) {
    OHS_CHECK_HANDLE(h);
    if (hrir_set < 0 || hrir_set >= h->cfg.n_hrir_sets) return fail(OHS_ERR_INVALID, "hrir_set %d out of range", hrir_set);
    if (path < 0 || path > 3) return fail(OHS_ERR_INVALID, "path %d out of range", path);
    if (len > 0 && !ir) return fail(OHS_ERR_INVALID, "null impulse response");
    if (len > (size_t)h->pmax * h->B) return fail(OHS_ERR_INVALID, "impulse response of %zu taps exceeds max_taps capacity %d", len, h->pmax * h->B);
    std::vector<float>& dst = h->h_ir[(size_t)hrir_set * 4 + path];
    dst.assign(ir, ir + len);
    // ir.chunks(BLOCK_SIZE).count(), or one silent partition for an empty slice (src/dsp/convolution.rs:114-132)
    h->h_path_parts[(size_t)hrir_set * 4 + path] = len == 0 ? 1 : (int)((len + h->B - 1) / h->B);
    h->set_dirty[hrir_set] = 1;
    h->set_external[hrir_set] = 0;
    h->any_set_dirty = true;
    return OHS_OK;
}

int ohs_num_partitions(ohs_engine* h, int hrir_set, int path, int* out) {
    OHS_CHECK_HANDLE(h);
    if (!out || hrir_set < 0 || hrir_set >= h->cfg.n_hrir_sets || path < 0 || path > 3) return fail(OHS_ERR_INVALID, "bad argument");
    *out = h->h_path_parts[(size_t)hrir_set * 4 + path];
    return OHS_OK;
}

int ohs_bind_stream_hrir(ohs_engine* h, int stream, int hrir_set) {
    OHS_CHECK_HANDLE(h);
    if (hrir_set < 0 || hrir_set >= h->cfg.n_hrir_sets) return fail(OHS_ERR_INVALID, "hrir_set %d out of range", hrir_set);
    if (stream != OHS_ALL && (stream < 0 || stream >= h->cfg.n_streams)) return fail(OHS_ERR_INVALID, "stream %d out of range", stream);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    if (stream == OHS_ALL) std::fill(h->h_stream_hrir.begin(), h->h_stream_hrir.end(), hrir_set);
    else h->h_stream_hrir[stream] = hrir_set;
    h->bind_dirty = true;
    // a re-bound stream starts from empty history, like a ConvolutionEngine that was just given its IRs
    if (stream == OHS_ALL) {
        int rc = upload_bindings(h);
        if (rc) return rc;
        return clear_history(h, false);
    }
    { const int rc = ensure_ring(h); if (rc) return rc; }
    h->zlin_valid = false;
    OHS_CUDA(cudaMemsetAsync(h->d_fdl + (size_t)stream * h->pmax * h->N, 0, sizeof(float2) * (size_t)h->pmax * h->N, h->stream));
    OHS_CUDA(cudaMemsetAsync(h->d_prev + (size_t)stream * h->B, 0, sizeof(float2) * (size_t)h->B, h->stream));
    return OHS_OK;
}

int ohs_commit_filters(ohs_engine* h) {
    OHS_CHECK_HANDLE(h);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    return commit_filters(h);
}

int ohs_filter_table(ohs_engine* h, void** dev_ptr, size_t* bytes) {
    OHS_CHECK_HANDLE(h);
    if (!dev_ptr || !bytes) return fail(OHS_ERR_INVALID, "null output");
    *dev_ptr = h->d_filt;
    *bytes = h->filt_bytes;
    return OHS_OK;
}

int ohs_mark_filters_external(ohs_engine* h, int hrir_set, int partitions) {
    OHS_CHECK_HANDLE(h);
    if (hrir_set < 0 || hrir_set >= h->cfg.n_hrir_sets) return fail(OHS_ERR_INVALID, "hrir_set %d out of range", hrir_set);
    if (partitions < 1 || partitions > h->pmax) return fail(OHS_ERR_INVALID, "partitions %d out of range 1..%d", partitions, h->pmax);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    h->set_dirty[hrir_set] = 0;
    h->set_external[hrir_set] = 1;
    h->h_set_parts[hrir_set] = partitions;
    for (int p = 0; p < 4; ++p) h->h_path_parts[(size_t)hrir_set * 4 + p] = partitions;
    h->any_set_dirty = false;
    for (unsigned char d : h->set_dirty) h->any_set_dirty |= d != 0;
    OHS_CUDA(cudaMemcpyAsync(h->d_set_parts, h->h_set_parts.data(), sizeof(int) * h->cfg.n_hrir_sets, cudaMemcpyHostToDevice, h->stream));
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    return OHS_OK;
}

// ---- EQ -----------------------------------------------------------------------------------------------------
int ohs_eq_update_band(ohs_engine* h, int eq_set, int band, int filter_type, float fc, float q, float gain_db, int enabled) {
    OHS_CHECK_HANDLE(h);
    if (band < 0 || band >= h->cfg.n_bands) {
        if (eq_set < 0 || eq_set >= h->cfg.n_eq_sets) return fail(OHS_ERR_INVALID, "eq_set %d out of range", eq_set);
        return OHS_OK;
    }
    float c[5];
    int rc = eq_design_impl(filter_type, h->cfg.sample_rate, fc, q, gain_db, c);
    if (rc) return rc;
    return set_band_host(h, eq_set, band, c, enabled);
}

int ohs_eq_set_band(ohs_engine* h, int eq_set, int band, const float coeffs[5], int enabled) {
    OHS_CHECK_HANDLE(h);
    if (!coeffs) return fail(OHS_ERR_INVALID, "null coefficients");
    return set_band_host(h, eq_set, band, coeffs, enabled);
}

int ohs_bind_stream_eq(ohs_engine* h, int stream, int eq_set) {
    OHS_CHECK_HANDLE(h);
    if (eq_set < 0 || eq_set >= h->cfg.n_eq_sets) return fail(OHS_ERR_INVALID, "eq_set %d out of range", eq_set);
    if (stream != OHS_ALL && (stream < 0 || stream >= h->cfg.n_streams)) return fail(OHS_ERR_INVALID, "stream %d out of range", stream);
    if (stream == OHS_ALL) std::fill(h->h_stream_eq.begin(), h->h_stream_eq.end(), eq_set);
    else h->h_stream_eq[stream] = eq_set;
    h->bind_dirty = true;
    return OHS_OK;
}

int ohs_eq_reset(ohs_engine* h) {
    OHS_CHECK_HANDLE(h);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    OHS_CUDA(cudaMemsetAsync(h->d_eqs, 0, sizeof(float4) * (size_t)h->cfg.n_streams * kMaxBands, h->stream));
    return OHS_OK;
}

int ohs_eq_frequency_response(ohs_engine* h, int eq_set, float sample_rate, const float* freqs, float* out, size_t n) {
    OHS_CHECK_HANDLE(h);
    if (eq_set < 0 || eq_set >= h->cfg.n_eq_sets || !freqs || !out) return fail(OHS_ERR_INVALID, "bad argument");
    const float fs = sample_rate > 0.0f ? sample_rate : h->cfg.sample_rate;
    for (size_t i = 0; i < n; ++i) {
        float rr = 1.0f, ri = 0.0f;
        for (int b = 0; b < h->cfg.n_bands; ++b) {
            const float* c = h->h_eqc.data() + ((size_t)eq_set * kMaxBands + b) * kEqCoefStride;
            if (c[5] == 0.0f) continue;
            const float w = 2.0f * 3.14159265358979323846f * freqs[i] / fs;
            const float c1 = cosf(w), s1 = sinf(w), c2 = cosf(2.0f * w), s2 = sinf(2.0f * w);
            const float nr = c[0] + c[1] * c1 + c[2] * c2, ni = c[1] * s1 + c[2] * s2;
            const float dr = 1.0f + c[3] * c1 + c[4] * c2, di = c[3] * s1 + c[4] * s2;
            const float den = dr * dr + di * di;
            const float hr = (nr * dr + ni * di) / den, hi = (ni * dr - nr * di) / den;
            const float tr = rr * hr - ri * hi, ti = rr * hi + ri * hr;
            rr = tr; ri = ti;
        }
        out[i] = hypotf(rr, ri);
    }
    return OHS_OK;
}

// ---- switches -----------------------------------------------------------------------------------------------
int ohs_set_eq_enable(ohs_engine* h, int enable) { OHS_CHECK_HANDLE(h); h->eq_enable = enable ? 1 : 0; return OHS_OK; }
int ohs_set_conv_enable(ohs_engine* h, int enable) { OHS_CHECK_HANDLE(h); h->conv_enable = enable ? 1 : 0; return OHS_OK; }
int ohs_set_bypass(ohs_engine* h, int bypass) { OHS_CHECK_HANDLE(h); h->bypass = bypass ? 1 : 0; return OHS_OK; }

int ohs_set_gain(ohs_engine* h, int stream, float gain) {
    OHS_CHECK_HANDLE(h);
    if (stream != OHS_ALL && (stream < 0 || stream >= h->cfg.n_streams)) return fail(OHS_ERR_INVALID, "stream %d out of range", stream);
    if (stream == OHS_ALL) std::fill(h->h_gain.begin(), h->h_gain.end(), gain);
    else h->h_gain[stream] = gain;
    h->gain_dirty = true;
    return OHS_OK;
}

int ohs_conv_reset(ohs_engine* h) {
    OHS_CHECK_HANDLE(h);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    h->head = 0;
    return clear_history(h, false);
}

int ohs_set_time_batch(ohs_engine* h, int enable) { OHS_CHECK_HANDLE(h); h->time_batch = enable ? 1 : 0; return OHS_OK; }

int ohs_streams_per_cta(ohs_engine* h, int* out) {
    OHS_CHECK_HANDLE(h);
    if (!out) return fail(OHS_ERR_INVALID, "null output");
    *out = h->G;
    return OHS_OK;
}

int ohs_prepare(ohs_engine* h, size_t n_frames, int host_io) {
    OHS_CHECK_HANDLE(h);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = commit_filters(h);
    if (rc) return rc;
    size_t per_launch = n_frames;
    if (host_io) {
        per_launch = stage_chunk_frames(h, n_frames);
        rc = ensure_stage_buffers(h, per_launch);
        if (rc) return rc;
    }
    const int n_blocks = (int)((per_launch + h->B - 1) / h->B);
    if (time_batch_eligible(h, n_blocks)) {
        rc = ensure_time_batch_scratch(h, n_blocks);
        if (rc < 0) return rc;
    }
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    return OHS_OK;
}

int ohs_debug_trace(ohs_engine* h, unsigned long long* d_stamps) {
    OHS_CHECK_HANDLE(h);
#ifdef OHS_TRACE
    h->d_trace = d_stamps;
    return OHS_OK;
#else
    (void)d_stamps;
    return fail(OHS_ERR_INVALID, "this library was built without -DOHS_TRACE");
#endif
}

// ---- processing ---------------------------------------------------------------------------------------------
int ohs_process_device(ohs_engine* h, const float* d_in, float* d_out, size_t n_frames, size_t row_stride) {
    OHS_CHECK_HANDLE(h);
    int rc = check_audio_args(h, d_in, d_out, n_frames, row_stride, true);
    if (rc) return rc;
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    if (n_frames == 0) return OHS_OK;
    if (h->bypass) {  // master_bypass: buffer untouched, no state advances (src/lib.rs:1169)
        if (d_in != d_out)
            OHS_CUDA(cudaMemcpy2DAsync(d_out, row_stride * sizeof(float), d_in, row_stride * sizeof(float), n_frames * sizeof(float),
                                       (size_t)h->cfg.n_streams * 2, cudaMemcpyDeviceToDevice, h->stream));
        return OHS_OK;
    }
    rc = commit_filters(h);
    if (rc) return rc;
    RenderParams p{};
    p.in = d_in; p.out = d_out; p.row_stride = p.out_row_stride = (long long)row_stride;
    p.n_blocks = (int)((n_frames + h->B - 1) / h->B);
    p.tail_frames = (int)(n_frames - (size_t)(p.n_blocks - 1) * h->B);
    p.n_streams = h->cfg.n_streams;
    p.stream_hrir = h->d_stream_hrir; p.stream_eq = h->d_stream_eq; p.stream_gain = h->d_stream_gain;
    p.filt = h->d_filt; p.set_parts = h->d_set_parts; p.fdl = h->d_fdl; p.prev = h->d_prev;
    p.eqc = h->d_eqc; p.eqs = h->d_eqs; p.tw = h->d_tw;
    p.pmax = h->pmax; p.head = h->head; p.n_bands = h->cfg.n_bands;
    p.eq_enable = h->eq_enable && h->cfg.n_bands > 0; p.conv_enable = h->conv_enable;
    p.uniform_set = (h->cfg.n_hrir_sets == 1) ? 1 : 0;
    // one HRIR set shared by every stream and small enough: the kernel keeps its spectra in shared memory
    p.filt_in_smem = 0;
    if (h->cfg.n_hrir_sets == 1 && h->N <= 512 && (size_t)h->h_set_parts[0] * h->N * sizeof(float4) <= 16 * 1024)
        p.filt_in_smem = h->h_set_parts[0];
    // (off by default: an event between two launches keeps the second from starting while the first drains)
    if (h->timing) OHS_CUDA(cudaEventRecord(h->ev_k0, h->stream));
    // long responses over many blocks: convolve along time per bin instead of re-reading the delay line every block
    bool batched = time_batch_eligible(h, p.n_blocks);
    if (batched) {
        rc = process_time_batched(h, p, d_in, d_out, row_stride);
        if (rc < 0) return rc;
        batched = (rc == 0);
    }
    if (!batched) {
        if (h->conv_enable) { rc = ensure_ring(h); if (rc) return rc; }   // the kernel reads and extends the delay-line ring
        rc = launch_render(h, p);
        if (rc) return rc;
        if (h->conv_enable) {
            h->head = (int)(((size_t)h->head + p.n_blocks) % (size_t)h->pmax);
            h->zlin_valid = false;
        }
    }
    if (h->timing) { OHS_CUDA(cudaEventRecord(h->ev_k1, h->stream)); h->timed = true; }
    return OHS_OK;
}

int ohs_sync(ohs_engine* h) {
    OHS_CHECK_HANDLE(h);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    return OHS_OK;
}

int ohs_cuda_stream(ohs_engine* h, void** stream) {
    OHS_CHECK_HANDLE(h);
    if (!stream) return fail(OHS_ERR_INVALID, "null output");
    *stream = (void*)h->stream;
    return OHS_OK;
}

int ohs_launch_count(ohs_engine* h, uint64_t* out) {
    OHS_CHECK_HANDLE(h);
    if (!out) return fail(OHS_ERR_INVALID, "null output");
    *out = h->launches;
    return OHS_OK;
}

int ohs_enable_timing(ohs_engine* h, int enable) { OHS_CHECK_HANDLE(h); h->timing = enable != 0; if (!enable) h->timed = false; return OHS_OK; }

int ohs_last_kernel_ms(ohs_engine* h, float* ms) {
    OHS_CHECK_HANDLE(h);
    if (!ms) return fail(OHS_ERR_INVALID, "null output");
    if (!h->timed) return fail(OHS_ERR_INVALID, "no timed process call yet (ohs_enable_timing first)");
    OHS_CUDA(cudaEventSynchronize(h->ev_k1));
    OHS_CUDA(cudaEventElapsedTime(ms, h->ev_k0, h->ev_k1));
    return OHS_OK;
}

int ohs_mix_device(ohs_engine* h, const float* d_in, float* d_bus, size_t n_frames, size_t row_stride, size_t bus_stride) {
    OHS_CHECK_HANDLE(h);
    if (!d_in || !d_bus) return fail(OHS_ERR_INVALID, "null audio pointer");
    if (n_frames % 4 || row_stride % 4 || bus_stride % 4 || (((uintptr_t)d_in | (uintptr_t)d_bus) & 15u))
        return fail(OHS_ERR_ALIGNMENT, "mixdown needs 16-byte aligned pointers and frame counts / strides that are multiples of 4");
    if (row_stride < n_frames || bus_stride < n_frames) return fail(OHS_ERR_INVALID, "stride smaller than n_frames");
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    if (n_frames == 0) return OHS_OK;
    const unsigned threads = 256;
    dim3 grid((unsigned)((n_frames / 4 + threads - 1) / threads), 2);
    mix_streams_kernel<<<grid, threads, 0, h->stream>>>(d_in, d_bus, h->cfg.n_streams, n_frames, row_stride, bus_stride);
    OHS_CUDA(cudaGetLastError());
    h->launches++;
    return OHS_OK;
}

// ---- collectives ----------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi* nccl_api() {
    static NcclApi api;
    static std::atomic<int> state{0};   // 0 untried, 1 ready, -1 failed
    static std::atomic_flag busy = ATOMIC_FLAG_INIT;
    if (state.load(std::memory_order_acquire) == 0) {
        while (busy.test_and_set(std::memory_order_acquire)) {}
        if (state.load(std::memory_order_relaxed) == 0) {
            // a process that already holds an NCCL (e.g. torch's bundled one) gets that one back by SONAME
            void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            bool ok = lib != nullptr;
            auto sym = [&](const char* name) { void* p = lib ? dlsym(lib, name) : nullptr; ok = ok && p; return p; };
            api.lib = lib;
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
            api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
            state.store(ok ? 1 : -1, std::memory_order_release);
        }
        busy.clear(std::memory_order_release);
    }
    return state.load(std::memory_order_acquire) == 1 ? &api : nullptr;
}

#define OHS_NCCL(api, expr)                                                                                          \
    do {                                                                                                             \
        ncclResult_t _r = (expr);                                                                                    \
        if (_r != ncclSuccess) return fail(OHS_ERR_NCCL, "%s failed: %s", #expr, (api)->GetErrorString(_r));          \
    } while (0)
}  // namespace

struct ohs_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
};

int ohs_comm_unique_id(void* id128) {
    static_assert(sizeof(ncclUniqueId) == OHS_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id128) return fail(OHS_ERR_INVALID, "null output");
    const NcclApi* api = nccl_api();
    if (!api) return fail(OHS_ERR_NCCL, "libnccl.so.2 could not be opened: %s", dlerror() ? dlerror() : "symbol missing");
    ncclUniqueId id;
    OHS_NCCL(api, api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return OHS_OK;
}

int ohs_comm_create(ohs_comm** out, int world, int rank, int device, const void* id128) {
    if (!out || !id128) return fail(OHS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(OHS_ERR_INVALID, "rank %d outside world %d", rank, world);
    const NcclApi* api = nccl_api();
    if (!api) return fail(OHS_ERR_NCCL, "libnccl.so.2 could not be opened");
    OHS_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ohs_comm* c = new ohs_comm();
    c->world = world; c->rank = rank; c->device = device;
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; return fail(OHS_ERR_NCCL, "ncclCommInitRank failed: %s", api->GetErrorString(r)); }
    *out = c;
    return OHS_OK;
}

int ohs_comm_destroy(ohs_comm* c) {
    if (!c) return OHS_OK;
    const NcclApi* api = nccl_api();
    if (api && c->comm) { cudaSetDevice(c->device); api->CommDestroy(c->comm); }
    delete c;
    return OHS_OK;
}

int ohs_broadcast_hrir(ohs_engine* h, ohs_comm* c, int root) {
    OHS_CHECK_HANDLE(h);
    if (!c) return fail(OHS_ERR_INVALID, "null communicator");
    if (root < 0 || root >= c->world) return fail(OHS_ERR_INVALID, "root %d outside world %d", root, c->world);
    if (c->device != h->cfg.device) return fail(OHS_ERR_INVALID, "communicator is on device %d, engine on %d", c->device, h->cfg.device);
    const NcclApi* api = nccl_api();
    if (!api) return fail(OHS_ERR_NCCL, "libnccl.so.2 could not be opened");
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    const int n_sets = h->cfg.n_hrir_sets;
    if (c->rank == root) {
        int rc = commit_filters(h);
        if (rc) return rc;
    } else {
        int rc = upload_bindings(h);
        if (rc) return rc;
    }
    OHS_NCCL(api, api->GroupStart());
    OHS_NCCL(api, api->Broadcast(h->d_filt, h->d_filt, h->filt_bytes / sizeof(float), ncclFloat32, root, c->comm, h->stream));
    OHS_NCCL(api, api->Broadcast(h->d_set_parts, h->d_set_parts, (size_t)n_sets, ncclInt32, root, c->comm, h->stream));
    OHS_NCCL(api, api->GroupEnd());
    if (c->rank != root) {
        OHS_CUDA(cudaMemcpyAsync(h->h_set_parts.data(), h->d_set_parts, sizeof(int) * n_sets, cudaMemcpyDeviceToHost, h->stream));
        OHS_CUDA(cudaStreamSynchronize(h->stream));
        for (int s = 0; s < n_sets; ++s) {
            const int parts = h->h_set_parts[s];
            if (parts < 1 || parts > h->pmax) return fail(OHS_ERR_INVALID, "received partition count %d for HRIR set %d (capacity %d)", parts, s, h->pmax);
            for (int p = 0; p < 4; ++p) h->h_path_parts[(size_t)s * 4 + p] = parts;
            h->set_dirty[s] = 0;
            h->set_external[s] = 1;
        }
        h->any_set_dirty = false;
        // a freshly received response starts from empty history, like set_ir (src/dsp/convolution.rs:135-138)
        int rc = clear_history(h, false);
        if (rc) return rc;
        h->head = 0;
    }
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    return OHS_OK;
}

int ohs_reduce_bus(ohs_engine* h, ohs_comm* c, float* d_bus, size_t n_floats, int root) {
    OHS_CHECK_HANDLE(h);
    if (!c || !d_bus) return fail(OHS_ERR_INVALID, "null argument");
    if (root < 0 || root >= c->world) return fail(OHS_ERR_INVALID, "root %d outside world %d", root, c->world);
    const NcclApi* api = nccl_api();
    if (!api) return fail(OHS_ERR_NCCL, "libnccl.so.2 could not be opened");
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    OHS_NCCL(api, api->Reduce(d_bus, d_bus, n_floats, ncclFloat32, ncclSum, root, c->comm, h->stream));
    return OHS_OK;
}

int ohs_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(OHS_ERR_INVALID, "null output");
    OHS_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return OHS_OK;
}

int ohs_host_free(void* p) {
    if (p) OHS_CUDA(cudaFreeHost(p));
    return OHS_OK;
}

// Host-pointer flavour: time chunks are staged through three HBM buffers; copy-in, kernel and copy-out of
// consecutive chunks overlap on three CUDA streams.
int ohs_process(ohs_engine* h, const float* in, float* out, size_t n_frames, size_t row_stride) {
    OHS_CHECK_HANDLE(h);
    int rc = check_audio_args(h, in, out, n_frames, row_stride, false);
    if (rc) return rc;
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    if (n_frames == 0) return OHS_OK;
    const size_t rows = (size_t)h->cfg.n_streams * 2;
    if (h->bypass) {
        if (in != out) for (size_t r = 0; r < rows; ++r) memcpy(out + r * row_stride, in + r * row_stride, n_frames * sizeof(float));
        return OHS_OK;
    }
    // chunk: whole blocks, about 24 MiB per staging buffer, at least one block
    // 24 MiB by default: measured best on PCIe Gen5 x16 with both directions busy (tools/e2e_probe.py)
    const size_t chunk = stage_chunk_frames(h, n_frames);
    rc = ensure_stage_buffers(h, chunk);
    if (rc) return rc;
    rc = commit_filters(h);
    if (rc) return rc;
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    size_t done = 0;
    int idx = 0;
    bool used[kPipe] = {false, false, false};
    while (done < n_frames) {
        const size_t n = std::min(chunk, n_frames - done);
        const size_t pitch = (n + 3) / 4 * 4;  // device rows stay 16-byte aligned for a ragged EQ-only tail
        const int b = idx % kPipe;
        if (used[b]) OHS_CUDA(cudaStreamWaitEvent(h->h2d, h->ev_out[b], 0));  // buffer free once its copy-out finished
        OHS_CUDA(cudaMemcpy2DAsync(h->d_stage[b], pitch * sizeof(float), in + done, row_stride * sizeof(float), n * sizeof(float), rows,
                                   cudaMemcpyHostToDevice, h->h2d));
        OHS_CUDA(cudaEventRecord(h->ev_in[b], h->h2d));
        OHS_CUDA(cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
        rc = ohs_process_device(h, h->d_stage[b], h->d_stage[b], n, pitch);
        if (rc) return rc;
        OHS_CUDA(cudaEventRecord(h->ev_comp[b], h->stream));
        OHS_CUDA(cudaStreamWaitEvent(h->d2h, h->ev_comp[b], 0));
        OHS_CUDA(cudaMemcpy2DAsync(out + done, row_stride * sizeof(float), h->d_stage[b], pitch * sizeof(float), n * sizeof(float), rows,
                                   cudaMemcpyDeviceToHost, h->d2h));
        OHS_CUDA(cudaEventRecord(h->ev_out[b], h->d2h));
        used[b] = true;
        done += n;
        ++idx;
    }
    OHS_CUDA(cudaStreamSynchronize(h->d2h));
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    return OHS_OK;
}

// The reference's FIFO adaptation around whole engine blocks (src/dsp/convolution.rs:141-182): append the host block,
// run every complete engine block, hand back n frames if that many are ready, else silence (and keep what is queued).
int ohs_process_fifo(ohs_engine* h, const float* in, float* out, size_t n_frames, size_t row_stride) {
    OHS_CHECK_HANDLE(h);
    if (!in || !out) return fail(OHS_ERR_INVALID, "null audio pointer");
    if (row_stride < n_frames) return fail(OHS_ERR_INVALID, "row_stride %zu < n_frames %zu", row_stride, n_frames);
    const size_t rows = (size_t)h->cfg.n_streams * 2;
    if (h->bypass) {
        if (in != out) for (size_t r = 0; r < rows; ++r) memcpy(out + r * row_stride, in + r * row_stride, n_frames * sizeof(float));
        return OHS_OK;
    }
    const size_t need = std::max(h->fifo_in_len, h->fifo_out_len) + n_frames + (size_t)h->B;
    if (need > h->fifo_cap) {
        const size_t cap = std::max(need, h->fifo_cap * 2);
        std::vector<float> ni(rows * cap, 0.f), no(rows * cap, 0.f);
        for (size_t r = 0; r < rows; ++r) {
            if (h->fifo_in_len) memcpy(&ni[r * cap], &h->fifo_in[r * h->fifo_cap], h->fifo_in_len * sizeof(float));
            if (h->fifo_out_len) memcpy(&no[r * cap], &h->fifo_out[r * h->fifo_cap], h->fifo_out_len * sizeof(float));
        }
        h->fifo_in.swap(ni); h->fifo_out.swap(no); h->fifo_cap = cap;
    }
    const size_t cap = h->fifo_cap;
    for (size_t r = 0; r < rows; ++r) memcpy(&h->fifo_in[r * cap + h->fifo_in_len], in + r * row_stride, n_frames * sizeof(float));
    h->fifo_in_len += n_frames;
    const size_t whole = (h->fifo_in_len / h->B) * h->B;
    if (whole) {
        // EQ runs inside the engine on the same whole blocks; in the reference the EQ sees the host buffer before the
        // FIFO (src/lib.rs:1194-1199) — per-sample filtering, so the result is identical.
        std::vector<float> tmp(rows * whole);
        for (size_t r = 0; r < rows; ++r) memcpy(&tmp[r * whole], &h->fifo_in[r * cap], whole * sizeof(float));
        int rc = ohs_process(h, tmp.data(), tmp.data(), whole, whole);
        if (rc) return rc;
        for (size_t r = 0; r < rows; ++r) {
            memcpy(&h->fifo_out[r * cap + h->fifo_out_len], &tmp[r * whole], whole * sizeof(float));
            memmove(&h->fifo_in[r * cap], &h->fifo_in[r * cap + whole], (h->fifo_in_len - whole) * sizeof(float));
        }
        h->fifo_in_len -= whole;
        h->fifo_out_len += whole;
    }
    if (h->fifo_out_len >= n_frames) {
        for (size_t r = 0; r < rows; ++r) {
            memcpy(out + r * row_stride, &h->fifo_out[r * cap], n_frames * sizeof(float));
            memmove(&h->fifo_out[r * cap], &h->fifo_out[r * cap + n_frames], (h->fifo_out_len - n_frames) * sizeof(float));
        }
        h->fifo_out_len -= n_frames;
    } else {
        for (size_t r = 0; r < rows; ++r) memset(out + r * row_stride, 0, n_frames * sizeof(float));  // :176-181
    }
    return OHS_OK;
}

// ---- state export / import ----------------------------------------------------------------------------------
// blob = StateHeader | delay line | overlap-save blocks | biquad states | FIFO residue (input rows, output rows)
namespace {
struct StateHeader {
    uint32_t magic;      // 'OHSS'
    int32_t abi, n_streams, n_bands, pmax, block, head, reserved;
    uint64_t fifo_in_len, fifo_out_len, total_bytes;
};
constexpr uint32_t kStateMagic = 0x5353484Fu;
size_t state_need(const ohs_engine* h, size_t fifo_in, size_t fifo_out) {
    const size_t S = h->cfg.n_streams;
    return sizeof(StateHeader) + sizeof(float2) * S * h->pmax * h->N + sizeof(float2) * S * h->B + sizeof(float4) * S * kMaxBands +
           sizeof(float) * S * 2 * (fifo_in + fifo_out);
}
}  // namespace

int ohs_state_bytes(ohs_engine* h, size_t* bytes) {
    OHS_CHECK_HANDLE(h);
    if (!bytes) return fail(OHS_ERR_INVALID, "null output");
    *bytes = state_need(h, h->fifo_in_len, h->fifo_out_len);
    return OHS_OK;
}

int ohs_state_export(ohs_engine* h, void* host_buf, size_t bytes) {
    OHS_CHECK_HANDLE(h);
    const size_t need = state_need(h, h->fifo_in_len, h->fifo_out_len);
    if (!host_buf || bytes < need) return fail(OHS_ERR_INVALID, "state buffer too small (%zu < %zu)", bytes, need);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = commit_filters(h);
    if (rc) return rc;
    rc = ensure_ring(h);   // the blob carries the delay-line ring
    if (rc) return rc;
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    unsigned char* p = (unsigned char*)host_buf;
    const size_t S = h->cfg.n_streams;
    StateHeader hdr{kStateMagic, OHS_ABI_VERSION, h->cfg.n_streams, kMaxBands, h->pmax, h->B, h->head, 0,
                    (uint64_t)h->fifo_in_len, (uint64_t)h->fifo_out_len, (uint64_t)need};
    memcpy(p, &hdr, sizeof(hdr)); p += sizeof(hdr);
    const size_t n0 = sizeof(float2) * S * h->pmax * h->N, n1 = sizeof(float2) * S * h->B, n2 = sizeof(float4) * S * kMaxBands;
    OHS_CUDA(cudaMemcpy(p, h->d_fdl, n0, cudaMemcpyDeviceToHost)); p += n0;
    OHS_CUDA(cudaMemcpy(p, h->d_prev, n1, cudaMemcpyDeviceToHost)); p += n1;
    OHS_CUDA(cudaMemcpy(p, h->d_eqs, n2, cudaMemcpyDeviceToHost)); p += n2;
    // what ohs_process_fifo has queued (src/dsp/convolution.rs:73-76 input_buffer / output_buffer), row by row
    for (size_t r = 0; r < S * 2; ++r) { memcpy(p, h->fifo_in.data() + r * h->fifo_cap, h->fifo_in_len * sizeof(float)); p += h->fifo_in_len * sizeof(float); }
    for (size_t r = 0; r < S * 2; ++r) { memcpy(p, h->fifo_out.data() + r * h->fifo_cap, h->fifo_out_len * sizeof(float)); p += h->fifo_out_len * sizeof(float); }
    return OHS_OK;
}

int ohs_state_import(ohs_engine* h, const void* host_buf, size_t bytes) {
    OHS_CHECK_HANDLE(h);
    if (!host_buf || bytes < sizeof(StateHeader)) return fail(OHS_ERR_INVALID, "state blob too small for its header");
    const unsigned char* p = (const unsigned char*)host_buf;
    StateHeader hdr;
    memcpy(&hdr, p, sizeof(hdr)); p += sizeof(hdr);
    if (hdr.magic != kStateMagic || hdr.abi != OHS_ABI_VERSION) return fail(OHS_ERR_INVALID, "not a state blob of ABI version %d", OHS_ABI_VERSION);
    if (hdr.n_streams != h->cfg.n_streams || hdr.n_bands != kMaxBands || hdr.pmax != h->pmax || hdr.block != h->B)
        return fail(OHS_ERR_INVALID, "state blob geometry (%d streams, %d partitions, block %d) does not match this engine (%d, %d, %d)",
                    hdr.n_streams, hdr.pmax, hdr.block, h->cfg.n_streams, h->pmax, h->B);
    if (hdr.head < 0 || hdr.head >= h->pmax) return fail(OHS_ERR_INVALID, "state blob ring head %d outside 0..%d", hdr.head, h->pmax - 1);
    const size_t limit = (size_t)1 << 32;
    if (hdr.fifo_in_len > limit || hdr.fifo_out_len > limit) return fail(OHS_ERR_INVALID, "state blob FIFO lengths are implausible");
    const size_t need = state_need(h, (size_t)hdr.fifo_in_len, (size_t)hdr.fifo_out_len);
    if (hdr.total_bytes != need || bytes != need) return fail(OHS_ERR_INVALID, "state blob is %zu bytes, header says %llu, geometry needs %zu", bytes, (unsigned long long)hdr.total_bytes, need);
    OHS_CUDA(cudaSetDevice(h->cfg.device));
    int rc = commit_filters(h);  // pending set_ir would otherwise clear the imported history later
    if (rc) return rc;
    OHS_CUDA(cudaStreamSynchronize(h->stream));
    const size_t S = h->cfg.n_streams;
    const size_t n0 = sizeof(float2) * S * h->pmax * h->N, n1 = sizeof(float2) * S * h->B, n2 = sizeof(float4) * S * kMaxBands;
    h->head = hdr.head;
    h->ring_valid = true; h->zlin_valid = false;
    OHS_CUDA(cudaMemcpy(h->d_fdl, p, n0, cudaMemcpyHostToDevice)); p += n0;
    OHS_CUDA(cudaMemcpy(h->d_prev, p, n1, cudaMemcpyHostToDevice)); p += n1;
    OHS_CUDA(cudaMemcpy(h->d_eqs, p, n2, cudaMemcpyHostToDevice)); p += n2;
    const size_t fi = (size_t)hdr.fifo_in_len, fo = (size_t)hdr.fifo_out_len;
    const size_t cap = std::max(fi, fo) + (size_t)h->B;
    h->fifo_in.assign(S * 2 * cap, 0.f); h->fifo_out.assign(S * 2 * cap, 0.f); h->fifo_cap = cap;
    for (size_t r = 0; r < S * 2; ++r) { memcpy(h->fifo_in.data() + r * cap, p, fi * sizeof(float)); p += fi * sizeof(float); }
    for (size_t r = 0; r < S * 2; ++r) { memcpy(h->fifo_out.data() + r * cap, p, fo * sizeof(float)); p += fo * sizeof(float); }
    h->fifo_in_len = fi; h->fifo_out_len = fo;
    return OHS_OK;
}

}  // extern "C"
