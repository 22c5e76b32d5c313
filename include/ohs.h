/*
 * ohs.h — C ABI of the B200-native Open Headstage DSP engine (libohs_cuda.so).
 *
 * This is the drop-in boundary for ONE path of KiloHertzian/Open-Headstage: the DSP core that
 * `Plugin::process` runs per host buffer (reference src/lib.rs:1156-1211):
 *
 *     10-band parametric EQ  ->  4-path binaural partitioned convolution  ->  output gain
 *     (src/dsp/parametric_eq.rs)   (src/dsp/convolution.rs)                   (src/lib.rs:1202-1207)
 *
 * re-posed as a batched renderer: one engine handle owns `n_streams` independent stereo chains (each the
 * equivalent of one reference `ConvolutionEngine` + one `StereoParametricEQ`), resident in HBM, processed
 * together by hand-written sm_100a kernels.  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point cites the reference interface it replaces.  The Rust-side binding a maintainer of the
 * reference would add (extern "C" block, wrapper structs with the reference's method names, build.rs
 * additions) is in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 (OHS_OK) or a negative ohs_status; ohs_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing throws or aborts across the boundary (the reference
 *     panics on bad EQ parameters, src/dsp/parametric_eq.rs:111; here that is OHS_ERR_INVALID).
 *   - audio is planar f32: in/out[(stream*2 + channel) * row_stride + frame], channel 0 = left, 1 = right
 *     (reference: `let [left, right] = buffer.as_slice()`, src/lib.rs:1175).
 *   - one handle = one CUDA stream; a handle is not re-entrant (the reference API is `&mut self`); distinct
 *     handles are independent.
 *   - there is no CPU fallback: if no CUDA device is usable, ohs_create fails.
 *   - the environment is read once, in ohs_create: OHS_STREAMS_PER_CTA (1..7, which render-kernel instantiation the
 *     handle uses; default: chosen from n_streams and the SM count), OHS_TIME_BATCH (0 = ohs_set_time_batch(h, 0)),
 *     OHS_STAGE_MB (staging chunk of the host-pointer path, default 24), OHS_PDL (0 = no programmatic dependent launches),
 *     OHS_TB_OVERLAP / OHS_TB_CHUNK / OHS_TB_EQ_G / OHS_TB_EQ_SMEM_KB (A/B switches of the time-batched route).
 */
#ifndef OHS_H
#define OHS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OHS_ABI_VERSION 2

typedef enum ohs_status {
    OHS_OK = 0,
    OHS_ERR_INVALID = -1,     /* bad argument (null handle, index out of range, unsupported block size, ...) */
    OHS_ERR_CUDA = -2,        /* a CUDA runtime call failed; message holds cudaGetErrorString */
    OHS_ERR_NO_DEVICE = -3,   /* no usable CUDA device (no CPU fallback exists) */
    OHS_ERR_NYQUIST = -4,     /* EQ design: 2*fc > fs  (biquad Errors::OutsideNyquist) */
    OHS_ERR_NEGATIVE_Q = -5,  /* EQ design: q < 0      (biquad Errors::NegativeQ) */
    OHS_ERR_ALIGNMENT = -6,   /* device/host audio pointer or row stride not 16-byte aligned */
    OHS_ERR_NCCL = -7         /* libnccl could not be opened, or an NCCL call failed; message holds ncclGetErrorString */
} ohs_status;

/* src/dsp/convolution.rs:26-33  enum ConvolutionPath { Lsl, Lsr, Rsl, Rsr }
 * (L/R speaker -> l/r ear; out_l = Lsl + Rsl, out_r = Lsr + Rsr, :229-230) */
typedef enum ohs_path { OHS_PATH_LSL = 0, OHS_PATH_LSR = 1, OHS_PATH_RSL = 2, OHS_PATH_RSR = 3 } ohs_path;

/* src/dsp/parametric_eq.rs:25-35  enum FilterType (declaration order) */
typedef enum ohs_filter_type {
    OHS_FILTER_PEAK = 0, OHS_FILTER_LOWSHELF = 1, OHS_FILTER_HIGHSHELF = 2, OHS_FILTER_LOWPASS = 3,
    OHS_FILTER_HIGHPASS = 4, OHS_FILTER_BANDPASS = 5, OHS_FILTER_NOTCH = 6, OHS_FILTER_ALLPASS = 7
} ohs_filter_type;

#define OHS_ALL (-1)        /* "every stream" for the per-stream setters */
#define OHS_MAX_BANDS 10    /* NUM_EQ_BANDS, src/lib.rs:40 */

/* Replaces the compile-time constants and constructor arguments of the reference:
 *   BLOCK_SIZE / FFT_SIZE (src/dsp/convolution.rs:22-23), NUM_EQ_BANDS (src/lib.rs:40),
 *   StereoParametricEQ::new(num_bands, sample_rate) (src/dsp/parametric_eq.rs:132). */
typedef struct ohs_config {
    int32_t n_streams;    /* independent stereo chains held by this engine (>= 1) */
    int32_t block;        /* engine block B in frames: 64, 128, 256, 512 or 1024; FFT size is 2*B */
    int32_t max_taps;     /* longest impulse response set_ir will be given; sizes the delay line: P = ceil(max_taps/B) */
    int32_t n_bands;      /* EQ bands per channel, 0..OHS_MAX_BANDS */
    int32_t n_hrir_sets;  /* distinct 4-path HRIR sets (>= 1); every stream starts bound to set 0 */
    int32_t n_eq_sets;    /* distinct EQ coefficient sets (>= 1); every stream starts bound to set 0 */
    int32_t device;       /* CUDA device ordinal */
    float sample_rate;    /* Hz; used by ohs_eq_update_band */
} ohs_config;

typedef struct ohs_engine ohs_engine;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
/* ConvolutionEngine::new (src/dsp/convolution.rs:87-108) + StereoParametricEQ::new (src/dsp/parametric_eq.rs:132-142)
 * for n_streams chains: every path's IR is silence (one all-zero partition, :44-65), every band is
 * PeakingEQ(0 dB)@20 Hz Q 0.707 and disabled (:63-76), gain 1, eq_enable 0 (src/lib.rs:433-434), bypass 0. */
int ohs_create(const ohs_config* cfg, ohs_engine** out);
int ohs_destroy(ohs_engine* h);
int ohs_abi_version(void);
const char* ohs_last_error(void);

/* ---- HRIR set-up ------------------------------------------------------------------------------------------ */
/* ConvolutionEngine::set_ir(path, &[f32]) (src/dsp/convolution.rs:111-139).  The IR is cut into B-frame chunks,
 * each zero-padded to 2B and transformed on the GPU; len == 0 gives one silent partition (:114-118).
 * As in the reference (:135-138) the convolution history of every stream bound to `hrir_set` is cleared.
 * (Difference, documented in DESIGN.md: the reference clears only the history of the one path; here the delay
 * line is shared by the four paths, so all four restart.  Identical whenever IRs are set before streaming.) */
int ohs_set_ir(ohs_engine* h, int hrir_set, int path, const float* ir, size_t len);
/* ir_fft_partitions.len() of that path, as asserted by the reference test (src/dsp/convolution.rs:395-399). */
int ohs_num_partitions(ohs_engine* h, int hrir_set, int path, int* out);
/* Which HRIR set a stream (or OHS_ALL) convolves with; clears that stream's convolution history. */
int ohs_bind_stream_hrir(ohs_engine* h, int stream, int hrir_set);
/* Device-resident filter spectra of all sets, for a caller-side NCCL broadcast: pointer and byte
 * size of the table the render kernel reads.  ohs_commit_filters uploads pending set_ir work first; after an
 * external overwrite (broadcast receive) call ohs_mark_filters_external so pending host IRs do not overwrite it. */
int ohs_commit_filters(ohs_engine* h);
int ohs_filter_table(ohs_engine* h, void** dev_ptr, size_t* bytes);
int ohs_mark_filters_external(ohs_engine* h, int hrir_set, int partitions);

/* ---- EQ --------------------------------------------------------------------------------------------------- */
/* biquad 0.4.2 Coefficients::<f32>::from_params as called by BiquadFilter::update_coeffs
 * (src/dsp/parametric_eq.rs:86-114).  Pure host function; out = {b0, b1, b2, a1, a2} normalised by a0. */
int ohs_eq_design(int filter_type, float fs, float fc, float q, float gain_db, float out[5]);
/* StereoParametricEQ::update_band_coeffs(band_idx, sample_rate, &BandConfig) (src/dsp/parametric_eq.rs:144-164):
 * same coefficients and `enabled` for left and right, filter state kept.  band >= n_bands is ignored (returns OK)
 * exactly like the reference (:145).  Uses cfg.sample_rate. */
int ohs_eq_update_band(ohs_engine* h, int eq_set, int band, int filter_type, float fc, float q, float gain_db, int enabled);
/* Same, with the five coefficients supplied as data. */
int ohs_eq_set_band(ohs_engine* h, int eq_set, int band, const float coeffs[5], int enabled);
int ohs_bind_stream_eq(ohs_engine* h, int stream, int eq_set);
/* StereoParametricEQ::reset_all_bands_state (src/dsp/parametric_eq.rs:181-188; Plugin::reset src/lib.rs:1152-1154). */
int ohs_eq_reset(ohs_engine* h);
/* StereoParametricEQ::calculate_frequency_response(sample_rate, frequencies) (src/dsp/parametric_eq.rs:191-209):
 * |product over the enabled bands of `eq_set` of H(e^{j 2 pi f / sample_rate})|; sample_rate <= 0 means cfg.sample_rate. */
int ohs_eq_frequency_response(ohs_engine* h, int eq_set, float sample_rate, const float* freqs, float* out, size_t n);

/* ---- chain switches (Plugin::process, src/lib.rs:1169-1207) ------------------------------------------------- */
int ohs_set_eq_enable(ohs_engine* h, int enable);          /* params.eq_enable (:1179) */
int ohs_set_conv_enable(ohs_engine* h, int enable);        /* 0 = EQ/gain only: StereoParametricEQ::process_block alone */
int ohs_set_bypass(ohs_engine* h, int bypass);             /* params.master_bypass (:1169): output = input, state untouched */
int ohs_set_gain(ohs_engine* h, int stream, float gain);   /* output_gain, one value per call (:1202-1207); stream or OHS_ALL */
/* Clears convolution history (as a fresh ConvolutionEngine with the same IRs would have). */
int ohs_conv_reset(ohs_engine* h);
/* 0 keeps long responses on the block-by-block kernel (see ohs_process_device); default 1. */
int ohs_set_time_batch(ohs_engine* h, int enable);
/* Streams each CTA of the render kernel carries for this handle (which instantiation render_kernel<2*block, G> runs). */
int ohs_streams_per_cta(ohs_engine* h, int* out);
/* Set-up that would otherwise happen lazily inside the first process call of that size: uploads pending set_ir /
 * EQ / binding work and allocates the scratch (time-batched route) and, for host_io != 0, the staging buffers that
 * calls of n_frames frames need.  Optional; lets a caller keep allocation out of a timed or real-time region, the
 * way Plugin::initialize (src/lib.rs:1124-1150) allocates before process() runs. */
int ohs_prepare(ohs_engine* h, size_t n_frames, int host_io);

/* ---- processing ------------------------------------------------------------------------------------------- */
/* The per-block process call: EQ (if enabled) -> convolution -> gain for every stream, n_frames frames each
 * (ConvolutionEngine::process_block src/dsp/convolution.rs:141 + StereoParametricEQ::process_block
 * src/dsp/parametric_eq.rs:166 + gain loop src/lib.rs:1202-1207).  n_frames must be a multiple of B: whole engine
 * blocks, i.e. the zero-latency case of the reference's FIFO (host block = multiple of BLOCK_SIZE).
 * in == out (in place) is allowed.  row_stride = frames between consecutive (stream, channel) rows (>= n_frames).
 * Device flavour: pointers are device memory on cfg.device; the call only enqueues on the engine's stream.
 * Long responses (>= 8 partitions) rendered >= 8 blocks per call take a time-batched route (EQ pre-pass on a second
 * stream of the handle, one chunk ahead of: forward transforms, a per-bin convolution along time, inverse transforms;
 * up to 2 GiB of scratch, allocated on first use or by ohs_prepare) with the same results within round-off and the
 * same state afterwards; the call is still ordered on the engine's stream alone; ohs_set_time_batch(h, 0) disables
 * it. */
int ohs_process_device(ohs_engine* h, const float* d_in, float* d_out, size_t n_frames, size_t row_stride);
/* Host flavour: pointers are host memory (pinned memory from ohs_host_alloc gives full PCIe speed); the call
 * stages time chunks through HBM with copies overlapped against the kernels and returns when `out` is complete. */
int ohs_process(ohs_engine* h, const float* in, float* out, size_t n_frames, size_t row_stride);
/* Arbitrary host-block length for ONE-call-per-host-buffer use: the reference's input/output FIFO adaptation
 * (src/dsp/convolution.rs:141-182) including the zero-filled output while fewer than n frames are ready. */
int ohs_process_fifo(ohs_engine* h, const float* in, float* out, size_t n_frames, size_t row_stride);
int ohs_sync(ohs_engine* h);
/* The engine's cudaStream_t (as void*) so callers can order their own work / NCCL calls against it. */
int ohs_cuda_stream(ohs_engine* h, void** stream);
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches evidence). */
int ohs_launch_count(ohs_engine* h, uint64_t* out);
/* Elapsed device time (ms) of the kernels of the most recent ohs_process_device call made while timing was enabled;
 * blocks until they are done.  Timing is off by default: it brackets every call with two CUDA events, which keeps
 * back-to-back per-block calls from overlapping their launch with the previous call's tail. */
int ohs_enable_timing(ohs_engine* h, int enable);
int ohs_last_kernel_ms(ohs_engine* h, float* ms);

/* Object mixdown (BASELINE config 4; an extension — the reference has no multi-source mode): sums the rendered
 * streams of this engine into one stereo bus, d_bus[c * bus_stride + n] = sum_s d_in[(s*2 + c) * row_stride + n].
 * Two mono sources ride in one stereo stream: source A as the left input with paths (LSL, LSR) = its (left-ear,
 * right-ear) HRIRs, source B as the right input with (RSL, RSR); the engine's output is already their binaural mix
 * (src/dsp/convolution.rs:229-230).  n_frames must be a multiple of 4; pointers 16-byte aligned.  Enqueues on the
 * engine's stream.  The per-GPU buses are then summed with NCCL (parallel.reduce_bus). */
int ohs_mix_device(ohs_engine* h, const float* d_in, float* d_bus, size_t n_frames, size_t row_stride, size_t bus_stride);

/* ---- multi-GPU collectives -------------------------------------------------------------------------
 * Streams shard across GPUs with no data-path collective; NCCL is used for exactly two things, both available here so
 * that a host without Python has the N > 1 path: one HRIR spectra table for the whole job (set_ir runs on one rank
 * only) and, for the object mixdown (config 4), the sum of the per-GPU stereo buses.  libnccl.so.2 is opened at run
 * time (dlopen) by the first ohs_comm_* call; a process that never calls them does not need NCCL.
 * One rank obtains the 128-byte id (ncclGetUniqueId) and ships it to the others by any means (the reference has no
 * multi-process mode; under torchrun parallel.create_comm ships it with torch.distributed). */
typedef struct ohs_comm ohs_comm;
#define OHS_COMM_ID_BYTES 128
int ohs_comm_unique_id(void* id128);
int ohs_comm_create(ohs_comm** out, int world, int rank, int device, const void* id128);   /* ncclCommInitRank */
int ohs_comm_destroy(ohs_comm* c);
/* Rank `root` uploads its pending set_ir work and broadcasts the device-resident spectra table and the per-set partition
 * counts on the engine's stream; the other ranks' sets are marked as externally filled (ohs_mark_filters_external).
 * Returns after the stream has drained. */
int ohs_broadcast_hrir(ohs_engine* h, ohs_comm* c, int root);
/* In-place ncclReduce(sum) of d_bus[n_floats] onto rank `root`, enqueued on the engine's stream (after ohs_mix_device). */
int ohs_reduce_bus(ohs_engine* h, ohs_comm* c, float* d_bus, size_t n_floats, int root);

/* pinned host memory helpers */
int ohs_host_alloc(void** p, size_t bytes);
int ohs_host_free(void* p);

/* Debug builds only (-DOHS_TRACE): d_stamps[CTA][16] receives clock64 stamps of the render kernel's milestones
 * (tools/trace_k1.py); a regular build returns OHS_ERR_INVALID. */
int ohs_debug_trace(ohs_engine* h, unsigned long long* d_stamps);

/* ---- state export / import (chunked offline renders resume bit-exactly) ----------------------
 * The blob holds a self-describing header (geometry, ring head, byte count), the delay line, the overlap-save blocks,
 * the biquad states and whatever ohs_process_fifo has queued; import rejects a blob whose geometry, size or ring
 * head does not match this engine.  ohs_state_bytes is the size an export would need NOW (it grows with the FIFO). */
int ohs_state_bytes(ohs_engine* h, size_t* bytes);
int ohs_state_export(ohs_engine* h, void* host_buf, size_t bytes);
int ohs_state_import(ohs_engine* h, const void* host_buf, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* OHS_H */
