/*
 * ohs_oracle.c — CPU restatement ("ref32") of the Open Headstage DSP hot path.  See ohs_oracle.h for the
 * scope rules (TEST INFRASTRUCTURE ONLY), the third-party arithmetic this restates and the parity pinning.
 *
 * Structure deliberately follows the reference, including its redundancy: four independent paths, each
 * with its own forward FFT of the input block, its own full-spectrum history ring and its own inverse FFT
 * (src/dsp/convolution.rs:193-224), so the CPU baseline timed from this file costs what the reference's
 * algorithm costs.  Spectra are kept as split re[]/im[] arrays (the reference uses interleaved
 * Complex<f32>); the arithmetic per element is identical.
 */
#define _GNU_SOURCE
#include "ohs_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sched.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------
 * c32 FFT, unnormalised, forward = negative exponent (rustfft contract; src/dsp/convolution.rs:88-90).
 * Radix-2 decimation in time, twiddles computed in f64 and rounded once to f32 (as rustfft does).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int n, log2n;
    int* bitrev;      /* [n] */
    float* tw_re;     /* per-stage contiguous tables: stage with half-length h starts at offset h-1 */
    float* tw_im;     /* forward sign (negative imaginary) */
} fft_plan;

static fft_plan* fft_plan_new(int n) {
    fft_plan* p = (fft_plan*)calloc(1, sizeof(fft_plan));
    p->n = n;
    p->log2n = 0;
    while ((1 << p->log2n) < n) p->log2n++;
    p->bitrev = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int b = 0; b < p->log2n; ++b) r |= ((i >> b) & 1) << (p->log2n - 1 - b);
        p->bitrev[i] = r;
    }
    p->tw_re = (float*)malloc(sizeof(float) * (size_t)n);
    p->tw_im = (float*)malloc(sizeof(float) * (size_t)n);
    for (int h = 1; h < n; h <<= 1) {
        for (int j = 0; j < h; ++j) {
            double a = -M_PI * (double)j / (double)h; /* -2*pi*j/(2h) */
            p->tw_re[h - 1 + j] = (float)cos(a);
            p->tw_im[h - 1 + j] = (float)sin(a);
        }
    }
    return p;
}

static void fft_plan_free(fft_plan* p) {
    if (!p) return;
    free(p->bitrev); free(p->tw_re); free(p->tw_im); free(p);
}

/* in-place on split arrays; `scratch_*` hold the bit-reversed copy */
static void fft_run(const fft_plan* p, float* re, float* im, float* sre, float* sim, int inverse) {
    const int n = p->n;
    for (int i = 0; i < n; ++i) { sre[p->bitrev[i]] = re[i]; sim[p->bitrev[i]] = im[i]; }
    for (int h = 1; h < n; h <<= 1) {
        const float* wr = p->tw_re + (h - 1);
        const float* wi = p->tw_im + (h - 1);
        const float s = inverse ? -1.0f : 1.0f;
        for (int base = 0; base < n; base += 2 * h) {
            float* ar = sre + base; float* ai = sim + base;
            float* br = ar + h;     float* bi = ai + h;
            for (int j = 0; j < h; ++j) {
                const float c = wr[j], d = s * wi[j];
                const float tr = br[j] * c - bi[j] * d;
                const float ti = br[j] * d + bi[j] * c;
                const float xr = ar[j], xi = ai[j];
                ar[j] = xr + tr; ai[j] = xi + ti;
                br[j] = xr - tr; bi[j] = xi - ti;
            }
        }
    }
    memcpy(re, sre, sizeof(float) * (size_t)n);
    memcpy(im, sim, sizeof(float) * (size_t)n);
}

/* ------------------------------------------------------------------------------------------------
 * Convolution engine — src/dsp/convolution.rs
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int num_partitions;   /* ir_fft_partitions.len()  (:38) */
    float* ir_re;         /* [P][N] */
    float* ir_im;
    float* hist_re;       /* input_fft_history [P][N] (:39) */
    float* hist_im;
    int history_index;    /* (:40) */
    float* overlap;       /* [B] (:41) */
} conv_path;

typedef struct { float* data; size_t len, cap; } fifo;

struct oracle_conv {
    int block, fft_size;
    conv_path paths[4];
    fft_plan* plan;
    fifo in_l, in_r, out_l, out_r;             /* (:76-79) */
    float *buf_re, *buf_im, *acc_re, *acc_im;  /* input_fft_buffer, conv_accumulator (:82-83) */
    float *scr_re, *scr_im;
    float *tmp_out[4];
    float *chunk_l, *chunk_r;
};

static void fifo_push(fifo* f, const float* x, size_t n) {
    if (f->len + n > f->cap) {
        size_t cap = f->cap ? f->cap : 1024;
        while (cap < f->len + n) cap *= 2;
        f->data = (float*)realloc(f->data, cap * sizeof(float));
        f->cap = cap;
    }
    memcpy(f->data + f->len, x, n * sizeof(float));
    f->len += n;
}

static void fifo_drain(fifo* f, float* dst, size_t n) {
    memcpy(dst, f->data, n * sizeof(float));
    memmove(f->data, f->data + n, (f->len - n) * sizeof(float));
    f->len -= n;
}

static void path_alloc(conv_path* p, int parts, int n, int b) {
    free(p->ir_re); free(p->ir_im); free(p->hist_re); free(p->hist_im);
    p->num_partitions = parts;
    p->ir_re = (float*)calloc((size_t)parts * n, sizeof(float));
    p->ir_im = (float*)calloc((size_t)parts * n, sizeof(float));
    p->hist_re = (float*)calloc((size_t)parts * n, sizeof(float));
    p->hist_im = (float*)calloc((size_t)parts * n, sizeof(float));
    p->history_index = 0;
    if (!p->overlap) p->overlap = (float*)calloc((size_t)b, sizeof(float));
}

/* ConvolutionEngine::new (:87-108) + ConvolutionPathData::new (:44-65): the default IR is BLOCK_SIZE zeros
 * -> one all-zero partition (silence, not passthrough). */
oracle_conv* oracle_conv_new(int block) {
    if (block < 2 || (block & (block - 1))) return NULL;
    oracle_conv* e = (oracle_conv*)calloc(1, sizeof(oracle_conv));
    e->block = block;
    e->fft_size = 2 * block;
    e->plan = fft_plan_new(e->fft_size);
    const size_t n = (size_t)e->fft_size;
    e->buf_re = (float*)calloc(n, sizeof(float)); e->buf_im = (float*)calloc(n, sizeof(float));
    e->acc_re = (float*)calloc(n, sizeof(float)); e->acc_im = (float*)calloc(n, sizeof(float));
    e->scr_re = (float*)calloc(n, sizeof(float)); e->scr_im = (float*)calloc(n, sizeof(float));
    for (int i = 0; i < 4; ++i) {
        e->tmp_out[i] = (float*)calloc((size_t)block, sizeof(float));
        path_alloc(&e->paths[i], 1, e->fft_size, block); /* FFT of zeros is zeros */
    }
    e->chunk_l = (float*)calloc((size_t)block, sizeof(float));
    e->chunk_r = (float*)calloc((size_t)block, sizeof(float));
    return e;
}

void oracle_conv_free(oracle_conv* e) {
    if (!e) return;
    for (int i = 0; i < 4; ++i) {
        conv_path* p = &e->paths[i];
        free(p->ir_re); free(p->ir_im); free(p->hist_re); free(p->hist_im); free(p->overlap);
        free(e->tmp_out[i]);
    }
    fft_plan_free(e->plan);
    free(e->buf_re); free(e->buf_im); free(e->acc_re); free(e->acc_im); free(e->scr_re); free(e->scr_im);
    free(e->chunk_l); free(e->chunk_r);
    free(e->in_l.data); free(e->in_r.data); free(e->out_l.data); free(e->out_r.data);
    free(e);
}

/* ConvolutionEngine::set_ir (:111-139) */
int oracle_conv_set_ir(oracle_conv* e, int path, const float* ir, size_t len) {
    if (path < 0 || path > 3) return -1;
    conv_path* p = &e->paths[path];
    const int n = e->fft_size, b = e->block;
    if (len == 0) {
        /* :114-118 empty IR -> one silent partition */
        path_alloc(p, 1, n, b);
    } else {
        /* :120-132 ir.chunks(BLOCK_SIZE), each zero-padded to FFT_SIZE and transformed */
        const int parts = (int)((len + (size_t)b - 1) / (size_t)b);
        path_alloc(p, parts, n, b);
        for (int k = 0; k < parts; ++k) {
            float* re = p->ir_re + (size_t)k * n;
            float* im = p->ir_im + (size_t)k * n;
            const size_t off = (size_t)k * b;
            const size_t m = (len - off) < (size_t)b ? (len - off) : (size_t)b;
            memset(re, 0, sizeof(float) * (size_t)n);
            memset(im, 0, sizeof(float) * (size_t)n);
            memcpy(re, ir + off, sizeof(float) * m);
            fft_run(e->plan, re, im, e->scr_re, e->scr_im, 0);
        }
    }
    /* :135-138 history ring re-created as zeros (done by path_alloc), index 0, overlap zeroed */
    p->history_index = 0;
    memset(p->overlap, 0, sizeof(float) * (size_t)b);
    return p->num_partitions;
}

int oracle_conv_num_partitions(const oracle_conv* e, int path) {
    return (path < 0 || path > 3) ? -1 : e->paths[path].num_partitions;
}

/* convolve_path_partitioned (:236-289) */
static void convolve_path_partitioned(oracle_conv* e, const float* input, conv_path* p, float* output) {
    const int n = e->fft_size, b = e->block, parts = p->num_partitions;
    float* xr = e->buf_re; float* xi = e->buf_im;
    /* 1. pack + zero-pad + forward FFT (:245-255) */
    for (int i = 0; i < b; ++i) { xr[i] = input[i]; xi[i] = 0.0f; }
    for (int i = b; i < n; ++i) { xr[i] = 0.0f; xi[i] = 0.0f; }
    fft_run(e->plan, xr, xi, e->scr_re, e->scr_im, 0);
    /* 2. store in history (:258) */
    memcpy(p->hist_re + (size_t)p->history_index * n, xr, sizeof(float) * (size_t)n);
    memcpy(p->hist_im + (size_t)p->history_index * n, xi, sizeof(float) * (size_t)n);
    /* 3. acc = sum_i X[(idx + P - i) % P] * H[i] over all FFT_SIZE bins (:261-273) */
    float* ar = e->acc_re; float* ai = e->acc_im;
    memset(ar, 0, sizeof(float) * (size_t)n);
    memset(ai, 0, sizeof(float) * (size_t)n);
    for (int i = 0; i < parts; ++i) {
        const int h = (p->history_index + parts - i) % parts;
        const float* hr = p->hist_re + (size_t)h * n; const float* hi = p->hist_im + (size_t)h * n;
        const float* gr = p->ir_re + (size_t)i * n;   const float* gi = p->ir_im + (size_t)i * n;
        for (int j = 0; j < n; ++j) {
            /* num_complex Mul: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re); then AddAssign per component */
            const float pr = hr[j] * gr[j] - hi[j] * gi[j];
            const float pi = hr[j] * gi[j] + hi[j] * gr[j];
            ar[j] += pr; ai[j] += pi;
        }
    }
    /* 4. inverse FFT, unnormalised (:276) */
    fft_run(e->plan, ar, ai, e->scr_re, e->scr_im, 1);
    /* 5. scale by 1/FFT_SIZE, overlap-add, save tail (:279-284) */
    const float scale = 1.0f / (float)n;
    for (int i = 0; i < b; ++i) {
        output[i] = ar[i] * scale + p->overlap[i];
        p->overlap[i] = ar[i + b] * scale;
    }
    p->history_index = (p->history_index + 1) % parts; /* :286 */
}

/* process_internal_block (:184-234) */
static void process_internal_block(oracle_conv* e, const float* in_l, const float* in_r, float* out_l, float* out_r) {
    convolve_path_partitioned(e, in_l, &e->paths[ORACLE_PATH_LSL], e->tmp_out[0]);
    convolve_path_partitioned(e, in_l, &e->paths[ORACLE_PATH_LSR], e->tmp_out[1]);
    convolve_path_partitioned(e, in_r, &e->paths[ORACLE_PATH_RSL], e->tmp_out[2]);
    convolve_path_partitioned(e, in_r, &e->paths[ORACLE_PATH_RSR], e->tmp_out[3]);
    for (int i = 0; i < e->block; ++i) { /* :228-231 */
        out_l[i] = e->tmp_out[0][i] + e->tmp_out[2][i];
        out_r[i] = e->tmp_out[1][i] + e->tmp_out[3][i];
    }
}

/* ConvolutionEngine::process_block (:141-182) */
void oracle_conv_process_block(oracle_conv* e, const float* in_l, const float* in_r, float* out_l, float* out_r, size_t n) {
    const size_t b = (size_t)e->block;
    fifo_push(&e->in_l, in_l, n); /* :149-150 */
    fifo_push(&e->in_r, in_r, n);
    while (e->in_l.len >= b) {    /* :152-161 */
        fifo_drain(&e->in_l, e->chunk_l, b);
        fifo_drain(&e->in_r, e->chunk_r, b);
        float pl[b], pr[b];
        process_internal_block(e, e->chunk_l, e->chunk_r, pl, pr);
        fifo_push(&e->out_l, pl, b);
        fifo_push(&e->out_r, pr, b);
    }
    if (e->out_l.len >= n) {      /* :163-175 */
        fifo_drain(&e->out_l, out_l, n);
        fifo_drain(&e->out_r, out_r, n);
    } else {                      /* :176-181 starved: silence, FIFO not drained */
        memset(out_l, 0, n * sizeof(float));
        memset(out_r, 0, n * sizeof(float));
    }
}

/* ------------------------------------------------------------------------------------------------
 * Parametric EQ — src/dsp/parametric_eq.rs over biquad 0.4.2
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float b0, b1, b2, a1, a2; /* Coefficients<f32>, a0-normalised */
    float s1, s2;             /* DirectForm2Transposed state */
    int enabled;
} biquad_filter;

struct oracle_eq {
    int num_bands;
    biquad_filter* left;
    biquad_filter* right;
};

/* biquad 0.4.2 Coefficients::<f32>::from_params as reached from BiquadFilter::update_coeffs
 * (src/dsp/parametric_eq.rs:94-111).  All arithmetic f32, libm sinf/cosf/powf/sqrtf. */
int oracle_eq_design(int filter_type, float fs, float fc, float q, float gain_db, float out[5]) {
    if (2.0f * fc > fs) return -1; /* Errors::OutsideNyquist */
    if (q < 0.0f) return -2;       /* Errors::NegativeQ */
    const float omega = 2.0f * (float)M_PI * fc / fs;
    const float omega_s = sinf(omega);
    const float omega_c = cosf(omega);
    const float alpha = omega_s / (2.0f * q);
    float b0, b1, b2, a0, a1, a2;
    int by_div = 0; /* shelves and peaking divide by a0; the others multiply by 1/a0 */
    switch (filter_type) {
    case ORACLE_FILTER_LOWPASS:
        b0 = (1.0f - omega_c) * 0.5f; b1 = 1.0f - omega_c; b2 = (1.0f - omega_c) * 0.5f;
        a0 = 1.0f + alpha; a1 = -2.0f * omega_c; a2 = 1.0f - alpha;
        break;
    case ORACLE_FILTER_HIGHPASS:
        b0 = (1.0f + omega_c) * 0.5f; b1 = -(1.0f + omega_c); b2 = (1.0f + omega_c) * 0.5f;
        a0 = 1.0f + alpha; a1 = -2.0f * omega_c; a2 = 1.0f - alpha;
        break;
    case ORACLE_FILTER_BANDPASS: /* constant skirt gain, peak gain = Q */
        b0 = omega_s / 2.0f; b1 = 0.0f; b2 = -(omega_s / 2.0f);
        a0 = 1.0f + alpha; a1 = -2.0f * omega_c; a2 = 1.0f - alpha;
        break;
    case ORACLE_FILTER_NOTCH:
        b0 = 1.0f; b1 = -2.0f * omega_c; b2 = 1.0f;
        a0 = 1.0f + alpha; a1 = -2.0f * omega_c; a2 = 1.0f - alpha;
        break;
    case ORACLE_FILTER_ALLPASS:
        b0 = 1.0f - alpha; b1 = -2.0f * omega_c; b2 = 1.0f + alpha;
        a0 = 1.0f + alpha; a1 = -2.0f * omega_c; a2 = 1.0f - alpha;
        break;
    case ORACLE_FILTER_LOWSHELF: {
        const float a = powf(10.0f, gain_db / 40.0f);
        const float sq = 2.0f * alpha * sqrtf(a);
        b0 = a * ((a + 1.0f) - (a - 1.0f) * omega_c + sq);
        b1 = 2.0f * a * ((a - 1.0f) - (a + 1.0f) * omega_c);
        b2 = a * ((a + 1.0f) - (a - 1.0f) * omega_c - sq);
        a0 = (a + 1.0f) + (a - 1.0f) * omega_c + sq;
        a1 = -2.0f * ((a - 1.0f) + (a + 1.0f) * omega_c);
        a2 = (a + 1.0f) + (a - 1.0f) * omega_c - sq;
        by_div = 1;
        break;
    }
    case ORACLE_FILTER_HIGHSHELF: {
        const float a = powf(10.0f, gain_db / 40.0f);
        const float sq = 2.0f * alpha * sqrtf(a);
        b0 = a * ((a + 1.0f) + (a - 1.0f) * omega_c + sq);
        b1 = -2.0f * a * ((a - 1.0f) + (a + 1.0f) * omega_c);
        b2 = a * ((a + 1.0f) + (a - 1.0f) * omega_c - sq);
        a0 = (a + 1.0f) - (a - 1.0f) * omega_c + sq;
        a1 = 2.0f * ((a - 1.0f) - (a + 1.0f) * omega_c);
        a2 = (a + 1.0f) - (a - 1.0f) * omega_c - sq;
        by_div = 1;
        break;
    }
    case ORACLE_FILTER_PEAK: {
        const float a = powf(10.0f, gain_db / 40.0f);
        b0 = 1.0f + alpha * a; b1 = -2.0f * omega_c; b2 = 1.0f - alpha * a;
        a0 = 1.0f + alpha / a; a1 = -2.0f * omega_c; a2 = 1.0f - alpha / a;
        by_div = 1;
        break;
    }
    default:
        return -3;
    }
    if (by_div) {
        out[0] = b0 / a0; out[1] = b1 / a0; out[2] = b2 / a0; out[3] = a1 / a0; out[4] = a2 / a0;
    } else {
        const float div = 1.0f / a0;
        out[0] = b0 * div; out[1] = b1 * div; out[2] = b2 * div; out[3] = a1 * div; out[4] = a2 * div;
    }
    return 0;
}

static void biquad_init(biquad_filter* f, float fs) {
    /* BiquadFilter::new (:63-76) */
    float c[5];
    oracle_eq_design(ORACLE_FILTER_PEAK, fs, 20.0f, 0.707f, 0.0f, c);
    f->b0 = c[0]; f->b1 = c[1]; f->b2 = c[2]; f->a1 = c[3]; f->a2 = c[4];
    f->s1 = 0.0f; f->s2 = 0.0f;
    f->enabled = 0;
}

oracle_eq* oracle_eq_new(int num_bands, float fs) {
    oracle_eq* q = (oracle_eq*)calloc(1, sizeof(oracle_eq));
    q->num_bands = num_bands;
    q->left = (biquad_filter*)calloc((size_t)num_bands, sizeof(biquad_filter));
    q->right = (biquad_filter*)calloc((size_t)num_bands, sizeof(biquad_filter));
    for (int i = 0; i < num_bands; ++i) { biquad_init(&q->left[i], fs); biquad_init(&q->right[i], fs); }
    return q;
}

void oracle_eq_free(oracle_eq* q) {
    if (!q) return;
    free(q->left); free(q->right); free(q);
}

void oracle_eq_set_band_raw(oracle_eq* q, int band_idx, const float c[5], int enabled) {
    if (band_idx < 0 || band_idx >= q->num_bands) return; /* :145 silently ignored */
    biquad_filter* f[2] = { &q->left[band_idx], &q->right[band_idx] };
    for (int k = 0; k < 2; ++k) {
        /* update_coefficients keeps s1/s2 (:112) */
        f[k]->b0 = c[0]; f[k]->b1 = c[1]; f[k]->b2 = c[2]; f[k]->a1 = c[3]; f[k]->a2 = c[4];
        f[k]->enabled = enabled ? 1 : 0; /* :153, :162 */
    }
}

int oracle_eq_update_band(oracle_eq* q, int band_idx, float fs, int filter_type, float fc, float qv, float gain_db, int enabled) {
    if (band_idx < 0 || band_idx >= q->num_bands) return 0;
    float c[5];
    const int rc = oracle_eq_design(filter_type, fs, fc, qv, gain_db, c);
    if (rc) return rc; /* the reference panics here (:111) */
    oracle_eq_set_band_raw(q, band_idx, c, enabled);
    return 0;
}

/* BiquadFilter::process_sample (:116-122) over DirectForm2Transposed::<f32>::run (biquad 0.4.2):
 *   out = s1 + b0*x;  s1 = s2 + b1*x - a1*out;  s2 = b2*x - a2*out        (each op separately rounded) */
static inline float biquad_process_sample(biquad_filter* f, float x) {
    if (!f->enabled) return x;
    const float out = f->s1 + f->b0 * x;
    f->s1 = f->s2 + f->b1 * x - f->a1 * out;
    f->s2 = f->b2 * x - f->a2 * out;
    return out;
}

void oracle_eq_process_block(oracle_eq* q, float* l, float* r, size_t n) {
    for (size_t i = 0; i < n; ++i) { /* :167-178 sample-outer, band-inner */
        float sl = l[i], sr = r[i];
        for (int j = 0; j < q->num_bands; ++j) {
            sl = biquad_process_sample(&q->left[j], sl);
            sr = biquad_process_sample(&q->right[j], sr);
        }
        l[i] = sl; r[i] = sr;
    }
}

void oracle_eq_reset(oracle_eq* q) {
    for (int i = 0; i < q->num_bands; ++i) {
        q->left[i].s1 = q->left[i].s2 = 0.0f;
        q->right[i].s1 = q->right[i].s2 = 0.0f;
    }
}

void oracle_eq_get_state(const oracle_eq* q, float* state) {
    for (int i = 0; i < q->num_bands; ++i) {
        state[i * 4 + 0] = q->left[i].s1;  state[i * 4 + 1] = q->left[i].s2;
        state[i * 4 + 2] = q->right[i].s1; state[i * 4 + 3] = q->right[i].s2;
    }
}

/* calculate_frequency_response (:191-209): z = from_polar(1, -omega); H = (b0 + b1 z + b2 z^2)/(1 + a1 z + a2 z^2)
 * (the reference writes z.powi(-1) with z = e^{-j omega}... it evaluates at z^-1 = e^{+j omega}; magnitude is the same). */
void oracle_eq_frequency_response(const oracle_eq* q, float fs, const float* freqs, float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float rr = 1.0f, ri = 0.0f;
        for (int j = 0; j < q->num_bands; ++j) {
            const biquad_filter* f = &q->left[j];
            if (!f->enabled) continue;
            const float omega = 2.0f * (float)M_PI * freqs[i] / fs;
            const float c1 = cosf(omega), s1 = sinf(omega);
            const float c2 = cosf(2.0f * omega), s2 = sinf(2.0f * omega);
            const float nr = f->b0 + f->b1 * c1 + f->b2 * c2, ni = f->b1 * s1 + f->b2 * s2;
            const float dr = 1.0f + f->a1 * c1 + f->a2 * c2, di = f->a1 * s1 + f->a2 * s2;
            const float den = dr * dr + di * di;
            const float hr = (nr * dr + ni * di) / den, hi = (ni * dr - nr * di) / den;
            const float tr = rr * hr - ri * hi, ti = rr * hi + ri * hr;
            rr = tr; ri = ti;
        }
        out[i] = hypotf(rr, ri);
    }
}

/* ------------------------------------------------------------------------------------------------
 * The chain — Plugin::process, src/lib.rs:1169-1207
 * ---------------------------------------------------------------------------------------------- */
void oracle_chain_process(oracle_conv* e, oracle_eq* q, int eq_enable, int bypass, float gain, float* l, float* r, size_t n) {
    if (bypass) return;                                      /* :1169 buffer untouched */
    if (eq_enable && q) oracle_eq_process_block(q, l, r, n); /* :1179-1195 */
    float* il = (float*)malloc(n * sizeof(float));           /* :1197-1198 to_vec */
    float* ir = (float*)malloc(n * sizeof(float));
    memcpy(il, l, n * sizeof(float));
    memcpy(ir, r, n * sizeof(float));
    oracle_conv_process_block(e, il, ir, l, r, n);           /* :1199-1200 */
    free(il); free(ir);
    for (size_t i = 0; i < n; ++i) { l[i] *= gain; r[i] *= gain; } /* :1202-1207 one gain per call */
}

/* ------------------------------------------------------------------------------------------------
 * Batched driver (CPU baseline): one stream per thread at a time, streams strided over the threads.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int tid, n_threads, n_streams, block, n_bands, eq_enable;
    const float* const* irs; const size_t* ir_len;
    const float* band_coeffs; const int* band_enabled;
    float gain;
    const float* in; float* out; size_t n_frames, host_block;
    double seconds;
} batch_job;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* batch_worker(void* arg) {
    batch_job* j = (batch_job*)arg;
#ifdef __linux__
    cpu_set_t set; CPU_ZERO(&set);
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (ncpu > 0) { CPU_SET(j->tid % ncpu, &set); pthread_setaffinity_np(pthread_self(), sizeof(set), &set); }
#endif
    double acc = 0.0;
    for (int s = j->tid; s < j->n_streams; s += j->n_threads) {
        oracle_conv* e = oracle_conv_new(j->block);
        oracle_eq* q = oracle_eq_new(j->n_bands, 48000.0f);
        for (int p = 0; p < 4; ++p) oracle_conv_set_ir(e, p, j->irs[p], j->ir_len[p]);
        for (int b = 0; b < j->n_bands; ++b) oracle_eq_set_band_raw(q, b, j->band_coeffs + 5 * b, j->band_enabled[b]);
        const float* il = j->in + ((size_t)s * 2 + 0) * j->n_frames;
        const float* ir = j->in + ((size_t)s * 2 + 1) * j->n_frames;
        float* ol = j->out + ((size_t)s * 2 + 0) * j->n_frames;
        float* orr = j->out + ((size_t)s * 2 + 1) * j->n_frames;
        memcpy(ol, il, j->n_frames * sizeof(float));
        memcpy(orr, ir, j->n_frames * sizeof(float));
        const double t0 = now_s();
        for (size_t off = 0; off < j->n_frames; off += j->host_block) {
            const size_t n = (j->n_frames - off) < j->host_block ? (j->n_frames - off) : j->host_block;
            oracle_chain_process(e, q, j->eq_enable, 0, j->gain, ol + off, orr + off, n);
        }
        acc += now_s() - t0;
        oracle_conv_free(e);
        oracle_eq_free(q);
    }
    j->seconds = acc;
    return NULL;
}

double oracle_render_batch(int n_streams, int n_threads, int block, const float* const irs[4], const size_t ir_len[4],
                           int n_bands, const float* band_coeffs, const int* band_enabled, int eq_enable,
                           float gain, const float* in, float* out, size_t n_frames, size_t host_block) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_streams) n_threads = n_streams;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    batch_job* jobs = (batch_job*)calloc((size_t)n_threads, sizeof(batch_job));
    const double t0 = now_s();
    for (int t = 0; t < n_threads; ++t) {
        batch_job* j = &jobs[t];
        j->tid = t; j->n_threads = n_threads; j->n_streams = n_streams; j->block = block; j->n_bands = n_bands;
        j->eq_enable = eq_enable; j->irs = irs; j->ir_len = ir_len; j->band_coeffs = band_coeffs;
        j->band_enabled = band_enabled; j->gain = gain; j->in = in; j->out = out; j->n_frames = n_frames;
        j->host_block = host_block;
        pthread_create(&th[t], NULL, batch_worker, j);
    }
    double worst = 0.0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].seconds > worst) worst = jobs[t].seconds;
    }
    const double wall = now_s() - t0;
    (void)wall;
    free(th); free(jobs);
    return worst; /* slowest thread's time inside its process loops */
}
