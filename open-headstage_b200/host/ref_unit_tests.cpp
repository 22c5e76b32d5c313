// The reference's five DSP unit tests (src/dsp/convolution.rs:317-421, src/dsp/parametric_eq.rs:218-238), written
// against the C++ mirror in dsp.hpp the way the reference writes them against its Rust objects.  Exit code 0 = pass.
#include <cmath>
#include <cstdio>
#include <vector>

#include "dsp.hpp"

using namespace ohs;

static const int BLOCK_SIZE = 512;   // src/dsp/convolution.rs:22
static const float TOLERANCE = 1e-3f;  // src/dsp/convolution.rs:301
static int failures = 0;

static void assert_approx_eq_slice(const float* a, const float* b, size_t n, float tol, const char* msg) {
    for (size_t i = 0; i < n; ++i)
        if (!(std::fabs(a[i] - b[i]) < tol)) {
            std::printf("FAIL %s: mismatch at %zu: %g vs %g\n", msg, i, a[i], b[i]);
            ++failures;
            return;
        }
    std::printf("ok   %s\n", msg);
}

static void test_identity_ir_passthrough() {
    ConvolutionEngine engine(BLOCK_SIZE);
    engine.set_ir(ConvolutionPath::Lsl, {1.0f});
    engine.set_ir(ConvolutionPath::Lsr, {0.0f});
    engine.set_ir(ConvolutionPath::Rsl, {0.0f});
    engine.set_ir(ConvolutionPath::Rsr, {1.0f});
    std::vector<float> in_l(BLOCK_SIZE), in_r(BLOCK_SIZE), out_l(BLOCK_SIZE), out_r(BLOCK_SIZE);
    for (int i = 0; i < BLOCK_SIZE; ++i) { in_l[i] = std::sin((float)i * 0.1f); in_r[i] = std::sin((float)i * -0.1f); }
    engine.process_block(in_l, in_r, out_l, out_r);
    engine.process_block(in_l, in_r, out_l, out_r);
    assert_approx_eq_slice(out_l.data(), in_l.data(), BLOCK_SIZE, TOLERANCE, "identity passthrough L");
    assert_approx_eq_slice(out_r.data(), in_r.data(), BLOCK_SIZE, TOLERANCE, "identity passthrough R");
}

static void test_delay_ir() {
    ConvolutionEngine engine(BLOCK_SIZE);
    const int delay = 5;
    std::vector<float> ir(delay + 1, 0.0f);
    ir[delay] = 1.0f;
    engine.set_ir(ConvolutionPath::Lsl, ir);
    engine.set_ir(ConvolutionPath::Lsr, {0.0f});
    engine.set_ir(ConvolutionPath::Rsl, {0.0f});
    engine.set_ir(ConvolutionPath::Rsr, {0.0f});
    const int n = BLOCK_SIZE * 2;
    std::vector<float> in_l(n), in_r(n, 0.0f), out_l(n), out_r(n), expected(n, 0.0f);
    for (int i = 0; i < n; ++i) in_l[i] = (float)i;
    engine.process_block(in_l, in_r, out_l, out_r);
    for (int i = delay; i < n; ++i) expected[i] = in_l[i - delay];
    assert_approx_eq_slice(out_l.data() + delay, expected.data() + delay, n - delay, TOLERANCE, "delayed signal");
}

static void test_long_ir_partitioning() {
    ConvolutionEngine engine(BLOCK_SIZE);
    const int ir_len = BLOCK_SIZE + BLOCK_SIZE / 2;
    std::vector<float> ir(ir_len, 0.0f);
    ir[0] = 1.0f;
    ir[ir_len - 1] = 0.5f;
    engine.set_ir(ConvolutionPath::Lsl, ir);
    if (engine.num_partitions(ConvolutionPath::Lsl) != 2) { std::printf("FAIL IR should be split into 2 partitions\n"); ++failures; }
    const int n = BLOCK_SIZE * 3;
    std::vector<float> in_l(n, 0.0f), in_r(n, 0.0f), out_l(n), out_r(n), expected(n, 0.0f);
    in_l[0] = 1.0f;
    engine.process_block(in_l, in_r, out_l, out_r);
    expected[0] = 1.0f;
    expected[ir_len - 1] = 0.5f;
    assert_approx_eq_slice(out_l.data(), expected.data(), ir_len, TOLERANCE, "long IR convolution");
}

static void test_biquad_filter_passthrough_when_disabled() {
    StereoParametricEQ eq(1, 48000.0f);  // every band starts disabled (src/dsp/parametric_eq.rs:74)
    std::vector<float> l{0.5f}, r{0.5f};
    eq.process_block(l, r);
    if (l[0] == 0.5f && r[0] == 0.5f) std::printf("ok   disabled filter is an exact passthrough\n");
    else { std::printf("FAIL disabled filter changed the sample: %g %g\n", l[0], r[0]); ++failures; }
}

static void test_biquad_filter_processes_when_enabled() {
    StereoParametricEQ eq(1, 48000.0f);
    BandConfig cfg;
    cfg.filter_type = FilterType::LowPass; cfg.center_freq = 1000.0f; cfg.q = 0.707f; cfg.gain_db = 0.0f; cfg.enabled = true;
    eq.update_band_coeffs(0, 48000.0f, cfg);
    std::vector<float> l{0.5f}, r{0.5f};
    eq.process_block(l, r);
    if (l[0] != 0.5f) std::printf("ok   enabled filter processes the sample (%g)\n", l[0]);
    else { std::printf("FAIL enabled filter left the sample unchanged\n"); ++failures; }
}

int main() {
    try {
        test_identity_ir_passthrough();
        test_delay_ir();
        test_long_ir_partitioning();
        test_biquad_filter_passthrough_when_disabled();
        test_biquad_filter_processes_when_enabled();
        // error behaviour: invalid EQ parameters are an error value, not a panic (src/dsp/parametric_eq.rs:111)
        StereoParametricEQ eq(10, 48000.0f);
        BandConfig bad;
        bad.center_freq = 30000.0f;
        bool threw = false;
        try { eq.update_band_coeffs(0, 48000.0f, bad); } catch (const Error& e) { threw = e.code == OHS_ERR_NYQUIST; }
        if (threw) std::printf("ok   fc above Nyquist is reported as OHS_ERR_NYQUIST\n"); else { std::printf("FAIL Nyquist check\n"); ++failures; }
    } catch (const std::exception& e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf("%s (%d failure(s))\n", failures ? "FAILED" : "PASSED", failures);
    return failures ? 1 : 0;
}
