// dsp.hpp — C++ mirror of the reference's two DSP objects over the C ABI of include/ohs.h.
//
// The reference is Rust (src/dsp/convolution.rs, src/dsp/parametric_eq.rs); no Rust toolchain exists in the build
// image, so the host side above the C ABI is written in C++ with the reference's type names, method names, argument
// meaning and error behaviour.  The Rust binding a maintainer would add is in INTEGRATION.md; it is the same thin
// layer in the other language.
//
//   ohs::ConvolutionEngine::{ConvolutionEngine, set_ir, process_block}      <- src/dsp/convolution.rs:87, 111, 141
//   ohs::StereoParametricEQ::{StereoParametricEQ, update_band_coeffs,
//                             process_block, reset_all_bands_state,
//                             calculate_frequency_response}                  <- src/dsp/parametric_eq.rs:132, 144, 166, 181, 191
//   ohs::BandConfig / ohs::FilterType / ohs::ConvolutionPath                 <- parametric_eq.rs:25-44, convolution.rs:26-33
//
// Everything computes on the GPU through libohs_cuda.so.  Errors surface as ohs::Error (the reference panics).
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ohs.h"

namespace ohs {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char* msg) : std::runtime_error(std::string("ohs: ") + msg), code(c) {}
};

inline void check(int rc) {
    if (rc != OHS_OK) throw Error(rc, ohs_last_error());
}

enum class ConvolutionPath { Lsl = OHS_PATH_LSL, Lsr = OHS_PATH_LSR, Rsl = OHS_PATH_RSL, Rsr = OHS_PATH_RSR };

enum class FilterType {
    Peak = OHS_FILTER_PEAK, LowShelf = OHS_FILTER_LOWSHELF, HighShelf = OHS_FILTER_HIGHSHELF, LowPass = OHS_FILTER_LOWPASS,
    HighPass = OHS_FILTER_HIGHPASS, BandPass = OHS_FILTER_BANDPASS, Notch = OHS_FILTER_NOTCH, AllPass = OHS_FILTER_ALLPASS
};

struct BandConfig {
    FilterType filter_type = FilterType::Peak;
    float center_freq = 1000.0f;
    float q = 0.707f;
    float gain_db = 0.0f;
    bool enabled = false;
};

namespace detail {
inline ohs_engine* make_engine(int n_streams, int block, int max_taps, int n_bands, float fs, int device) {
    ohs_config cfg{};
    cfg.n_streams = n_streams; cfg.block = block; cfg.max_taps = max_taps; cfg.n_bands = n_bands;
    cfg.n_hrir_sets = 1; cfg.n_eq_sets = 1; cfg.device = device; cfg.sample_rate = fs;
    ohs_engine* h = nullptr;
    check(ohs_create(&cfg, &h));
    return h;
}
}  // namespace detail

// One stereo stream.  BLOCK_SIZE (src/dsp/convolution.rs:22) is a constructor argument here.
class ConvolutionEngine {
  public:
    explicit ConvolutionEngine(int block_size = 512, int max_taps = 65536, int device = 0)
        : h_(detail::make_engine(1, block_size, max_taps, 0, 48000.0f, device)) {}
    ~ConvolutionEngine() { ohs_destroy(h_); }
    ConvolutionEngine(const ConvolutionEngine&) = delete;
    ConvolutionEngine& operator=(const ConvolutionEngine&) = delete;

    void set_ir(ConvolutionPath path, const std::vector<float>& ir_data) {
        check(ohs_set_ir(h_, 0, static_cast<int>(path), ir_data.data(), ir_data.size()));
    }
    int num_partitions(ConvolutionPath path) {
        int n = 0;
        check(ohs_num_partitions(h_, 0, static_cast<int>(path), &n));
        return n;
    }
    // Any host-block length; FIFO adaptation and zero-fill on starvation as in the reference (:141-182).
    void process_block(const std::vector<float>& input_left, const std::vector<float>& input_right, std::vector<float>& output_left,
                       std::vector<float>& output_right) {
        const size_t n = input_left.size();
        if (input_right.size() != n || output_left.size() != n || output_right.size() != n) throw Error(OHS_ERR_INVALID, "length mismatch");
        std::vector<float> io(2 * n);
        std::copy(input_left.begin(), input_left.end(), io.begin());
        std::copy(input_right.begin(), input_right.end(), io.begin() + n);
        check(ohs_process_fifo(h_, io.data(), io.data(), n, n));
        std::copy(io.begin(), io.begin() + n, output_left.begin());
        std::copy(io.begin() + n, io.end(), output_right.begin());
    }

  private:
    ohs_engine* h_;
};

class StereoParametricEQ {
  public:
    StereoParametricEQ(size_t num_bands, float initial_sample_rate, int device = 0)
        : h_(detail::make_engine(1, 256, 1, static_cast<int>(num_bands), initial_sample_rate, device)) {
        check(ohs_set_conv_enable(h_, 0));
        check(ohs_set_eq_enable(h_, 1));
    }
    ~StereoParametricEQ() { ohs_destroy(h_); }
    StereoParametricEQ(const StereoParametricEQ&) = delete;
    StereoParametricEQ& operator=(const StereoParametricEQ&) = delete;

    void update_band_coeffs(size_t band_idx, float sample_rate, const BandConfig& config) {
        float c[5];
        check(ohs_eq_design(static_cast<int>(config.filter_type), sample_rate, config.center_freq, config.q, config.gain_db, c));
        check(ohs_eq_set_band(h_, 0, static_cast<int>(band_idx), c, config.enabled ? 1 : 0));
    }
    // in place, like the reference (:166)
    void process_block(std::vector<float>& input_left, std::vector<float>& input_right) {
        const size_t n = input_left.size();
        if (input_right.size() != n) throw Error(OHS_ERR_INVALID, "length mismatch");
        std::vector<float> io(2 * n);
        std::copy(input_left.begin(), input_left.end(), io.begin());
        std::copy(input_right.begin(), input_right.end(), io.begin() + n);
        check(ohs_process(h_, io.data(), io.data(), n, n));
        std::copy(io.begin(), io.begin() + n, input_left.begin());
        std::copy(io.begin() + n, io.end(), input_right.begin());
    }
    void reset_all_bands_state() { check(ohs_eq_reset(h_)); }
    std::vector<float> calculate_frequency_response(float sample_rate, const std::vector<float>& frequencies) {
        std::vector<float> out(frequencies.size());
        check(ohs_eq_frequency_response(h_, 0, sample_rate, frequencies.data(), out.data(), out.size()));
        return out;
    }

  private:
    ohs_engine* h_;
};

}  // namespace ohs
