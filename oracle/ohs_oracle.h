/*
 * ohs_oracle.h — CPU restatement ("ref32") of the Open Headstage DSP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (open-headstage_b200/, include/, the C-ABI
 * library) may include, link or call this.  Allowed users: tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference repo root).
 * The reference cannot be compiled here (no cargo/rustc, no libmysofa) and its arithmetic lives in two
 * un-vendored crates, so this is a restatement, not a build of the reference:
 *   - rustfft 6.4.0 (Cargo.lock:2579-2591): unnormalised c32 FFT, forward = e^{-2*pi*i*jk/N}.  rustfft picks
 *     its algorithm by CPU feature at run time, so its rounding is machine-dependent; the contract with it
 *     is tolerance based.  Restated here as a plain radix-2 FFT with f64-computed twiddles.
 *   - biquad 0.4.2 (Cargo.lock:314-321): Coefficients::<f32>::from_params (RBJ cookbook) and
 *     DirectForm2Transposed::<f32>::run.  Restated from the crate's published source as recalled.
 *
 * PARITY PINNING.  Convolution: pinned against the reference's three known-answer unit tests
 * (src/dsp/convolution.rs:317-421) replayed verbatim in tests/test_oracle.py.  EQ: the reference holds no
 * value-pinning test for the biquad maths (src/dsp/parametric_eq.rs:218-238 pin only "disabled == exact
 * passthrough" and "enabled != input"), and the crate is absent, so EQ coefficient design and the DF2T op
 * order are "parity unpinned" against the real crate; they are pinned only against an independent f64
 * scipy.signal evaluation of the same cookbook formulae (tests/test_oracle.py).
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fno-fast-math (Rust never contracts a*b+c into an FMA;
 * the EQ output is bit-sensitive to contraction).
 */
#ifndef OHS_ORACLE_H
#define OHS_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/dsp/convolution.rs:26-33 — enum ConvolutionPath { Lsl, Lsr, Rsl, Rsr } */
enum { ORACLE_PATH_LSL = 0, ORACLE_PATH_LSR = 1, ORACLE_PATH_RSL = 2, ORACLE_PATH_RSR = 3 };

/* src/dsp/parametric_eq.rs:25-35 — enum FilterType (declaration order) */
enum {
    ORACLE_FILTER_PEAK = 0, ORACLE_FILTER_LOWSHELF = 1, ORACLE_FILTER_HIGHSHELF = 2, ORACLE_FILTER_LOWPASS = 3,
    ORACLE_FILTER_HIGHPASS = 4, ORACLE_FILTER_BANDPASS = 5, ORACLE_FILTER_NOTCH = 6, ORACLE_FILTER_ALLPASS = 7
};

/* ---- convolution engine: src/dsp/convolution.rs ---- */
typedef struct oracle_conv oracle_conv;
/* ConvolutionEngine::new (:87-108); `block` replaces the compile-time BLOCK_SIZE=512 (:22), FFT_SIZE = 2*block (:23). */
oracle_conv* oracle_conv_new(int block);
void oracle_conv_free(oracle_conv* e);
/* ConvolutionEngine::set_ir (:111-139).  Returns the number of partitions. */
int oracle_conv_set_ir(oracle_conv* e, int path, const float* ir, size_t len);
int oracle_conv_num_partitions(const oracle_conv* e, int path);
/* ConvolutionEngine::process_block (:141-182) including the host-block FIFO adaptation and the zero-fill on starvation. */
void oracle_conv_process_block(oracle_conv* e, const float* in_l, const float* in_r, float* out_l, float* out_r, size_t n);

/* ---- parametric EQ: src/dsp/parametric_eq.rs ---- */
typedef struct oracle_eq oracle_eq;
/* biquad 0.4.2 Coefficients::<f32>::from_params behind BiquadFilter::update_coeffs (:86-114).
 * out = {b0, b1, b2, a1, a2}.  Returns 0, or -1 (OutsideNyquist: 2*fc > fs) / -2 (NegativeQ) where the reference's
 * .unwrap() (:111) would panic. */
int oracle_eq_design(int filter_type, float fs, float fc, float q, float gain_db, float out[5]);
/* StereoParametricEQ::new (:132-142): every band = PeakingEQ(0 dB) @ 20 Hz, Q 0.707, disabled (:63-76). */
oracle_eq* oracle_eq_new(int num_bands, float fs);
void oracle_eq_free(oracle_eq* q);
/* StereoParametricEQ::update_band_coeffs (:144-164): same coefficients and `enabled` on L and R; state kept;
 * band_idx >= num_bands silently ignored.  Returns oracle_eq_design's code. */
int oracle_eq_update_band(oracle_eq* q, int band_idx, float fs, int filter_type, float fc, float qv, float gain_db, int enabled);
/* Same, with the five coefficients supplied as data (what the GPU engine ingests). */
void oracle_eq_set_band_raw(oracle_eq* q, int band_idx, const float coeffs[5], int enabled);
/* StereoParametricEQ::process_block (:166-179): in place, sample-outer / band-inner, DF2T without FMA. */
void oracle_eq_process_block(oracle_eq* q, float* l, float* r, size_t n);
/* StereoParametricEQ::reset_all_bands_state (:181-188) */
void oracle_eq_reset(oracle_eq* q);
/* StereoParametricEQ::calculate_frequency_response (:191-209), left bank, enabled bands only. */
void oracle_eq_frequency_response(const oracle_eq* q, float fs, const float* freqs, float* out, size_t n);
/* raw state access for state export/import tests: state = [band][channel(L,R)][s1,s2] */
void oracle_eq_get_state(const oracle_eq* q, float* state);

/* ---- the chain in Plugin::process: src/lib.rs:1169-1207 ---- */
/* bypass test -> EQ (if eq_enable) in place -> convolution (reads a copy, writes in place) -> every sample *= gain. */
void oracle_chain_process(oracle_conv* e, oracle_eq* q, int eq_enable, int bypass, float gain, float* l, float* r, size_t n);

/* ---- batched driver: one independent chain per stream, one stream per thread at a time ----
 * in/out laid out [stream][channel][n_frames]; every stream uses the same 4 IRs, band coefficients and gain
 * (BASELINE configs 2, 3, 5).  Feeds each chain host blocks of `host_block` frames.  Returns the wall-clock seconds
 * spent inside the per-stream process loops (engine construction and set_ir are outside the timed region). */
double oracle_render_batch(int n_streams, int n_threads, int block, const float* const irs[4], const size_t ir_len[4],
                           int n_bands, const float* band_coeffs /* [n_bands][5] */, const int* band_enabled, int eq_enable,
                           float gain, const float* in, float* out, size_t n_frames, size_t host_block);

#ifdef __cplusplus
}
#endif
#endif
