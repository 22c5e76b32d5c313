// sofa.hpp — host-side HRIR source for the engine in C++: the step immediately before `set_ir`.
//
// The reference does this in Rust over libmysofa (`MySofa::open` / `get_hrtf_irs`, src/sofa/loader.rs:79, 136) and — as
// shipped — never connects it to the convolver (SURVEY finding 5).  This is the equivalent a batch renderer needs,
// with the reference's names: `ohs::MySofa::open(path, target_samplerate)`, `filter_length()`,
// `get_hrtf_irs(azimuth_deg, elevation_deg, radius_m)`, plus `wire_speakers`, the SOFA -> four `set_ir` calls the
// reference intends (src/lib.rs:1136-1146, github_issues/sofa_implement_logic_select_extract_hrirs.md).
//
// No libmysofa and no HDF5 library: a SimpleFreeFieldHRIR file stores `SourcePosition[M][3]` and `Data.IR[M][2][N]` as
// f64 datasets that are zlib-compressed and byte-shuffled; they are located by scanning for zlib streams whose decoded
// size matches (zlib only), un-shuffled and narrowed to f32.  Selection is nearest neighbour on the unit sphere
// (bit-exact index lookup: tests/test_host_inputs.py); libmysofa's neighbour interpolation and resample-on-open are
// not reproduced — taps are used as measured, `target_samplerate` is recorded but does not resample.
//
// Conventions: SOFA spherical coordinates, azimuth in degrees counter-clockwise (positive = left), elevation in
// degrees.  The plugin's UI uses negative azimuth = left (src/lib.rs:429-431): ui_azimuth_to_sofa.
#pragma once

#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ohs {

struct SofaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace detail {

// first zlib stream in `buf` that inflates to exactly `decoded_size` bytes
inline bool find_zlib_stream(const std::vector<unsigned char>& buf, size_t decoded_size, std::vector<unsigned char>& out) {
    out.resize(decoded_size + 1);
    for (size_t i = 0; i + 1 < buf.size(); ++i) {
        // RFC 1950 header: 0x78 then 0x01 / 0x5E / 0x9C / 0xDA
        if (buf[i] != 0x78) continue;
        const unsigned char f = buf[i + 1];
        if (f != 0x01 && f != 0x5E && f != 0x9C && f != 0xDA) continue;
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK) throw SofaError("zlib inflateInit failed");
        zs.next_in = const_cast<unsigned char*>(buf.data() + i);
        zs.avail_in = static_cast<uInt>(std::min<size_t>(buf.size() - i, 0x7fffffffu));
        zs.next_out = out.data();
        zs.avail_out = static_cast<uInt>(out.size());   // one byte more than wanted: a longer stream is not a match
        const int rc = inflate(&zs, Z_FINISH);
        const size_t produced = zs.total_out;
        inflateEnd(&zs);
        if (rc == Z_STREAM_END && produced == decoded_size) { out.resize(decoded_size); return true; }
    }
    return false;
}

// HDF5 shuffle filter undone: byte plane b of element e sits at raw[b * count + e]
inline std::vector<double> unshuffle_f64(const std::vector<unsigned char>& raw) {
    const size_t count = raw.size() / 8;
    std::vector<double> v(count);
    for (size_t e = 0; e < count; ++e) {
        unsigned char bytes[8];
        for (int b = 0; b < 8; ++b) bytes[b] = raw[(size_t)b * count + e];
        std::memcpy(&v[e], bytes, 8);   // little-endian f64
    }
    return v;
}

}  // namespace detail

class MySofa {
  public:
    // [M][2][N] f32 and [M][3] (azimuth deg, elevation deg, radius m), e.g. from another reader
    static MySofa from_arrays(std::vector<float> ir, std::vector<float> position, size_t n_measurements, size_t n_taps, float sample_rate) {
        if (ir.size() != n_measurements * 2 * n_taps || position.size() != n_measurements * 3) throw SofaError("array sizes do not match M and N");
        MySofa s;
        s.ir_ = std::move(ir); s.pos_ = std::move(position); s.m_ = n_measurements; s.n_ = n_taps;
        s.source_samplerate_ = s.resampled_samplerate_ = sample_rate;
        return s;
    }

    // MySofa::open (src/sofa/loader.rs:79-134).  M and N are found by trying the common CIPIC/ARI/... shapes unless given.
    static MySofa open(const std::string& filepath, float target_samplerate, size_t n_measurements = 0, size_t n_taps = 0,
                       float source_samplerate = 44100.0f) {
        std::ifstream f(filepath, std::ios::binary);
        if (!f) throw SofaError("Failed to open SOFA file '" + filepath + "'");
        std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        static const size_t shapes[][2] = {{1250, 200}, {1550, 256}, {2304, 256}, {828, 256}, {710, 512}, {2702, 512}};
        std::vector<std::pair<size_t, size_t>> tries;
        if (n_measurements && n_taps) tries.emplace_back(n_measurements, n_taps);
        else for (const auto& s : shapes) tries.emplace_back(s[0], s[1]);
        for (const auto& mn : tries) {
            std::vector<unsigned char> raw_pos, raw_ir;
            if (!detail::find_zlib_stream(buf, mn.first * 3 * 8, raw_pos)) continue;
            if (!detail::find_zlib_stream(buf, mn.first * 2 * mn.second * 8, raw_ir)) continue;
            const std::vector<double> pos = detail::unshuffle_f64(raw_pos), ir = detail::unshuffle_f64(raw_ir);
            MySofa s;
            s.m_ = mn.first; s.n_ = mn.second;
            s.pos_.assign(pos.begin(), pos.end());
            s.ir_.assign(ir.begin(), ir.end());
            s.source_samplerate_ = source_samplerate;
            s.resampled_samplerate_ = target_samplerate;
            return s;
        }
        throw SofaError("could not locate SourcePosition / Data.IR datasets in '" + filepath + "'");
    }

    size_t filter_length() const { return n_; }
    size_t measurements() const { return m_; }
    float source_samplerate() const { return source_samplerate_; }
    float resampled_samplerate() const { return resampled_samplerate_; }
    const std::vector<float>& ir() const { return ir_; }
    const std::vector<float>& position() const { return pos_; }

    // index of the measurement closest (great circle) to the direction
    size_t nearest(float azimuth_deg, float elevation_deg) const {
        const double d2r = 3.14159265358979323846 / 180.0;
        double az = std::fmod((double)azimuth_deg, 360.0);
        if (az < 0) az += 360.0;
        const double a = az * d2r, e = (double)elevation_deg * d2r;
        size_t best = 0;
        double best_dot = -2.0;
        for (size_t i = 0; i < m_; ++i) {
            const double azi = (double)pos_[3 * i] * d2r, eli = (double)pos_[3 * i + 1] * d2r;
            const double dot = std::cos(eli) * std::cos(e) * std::cos(azi - a) + std::sin(eli) * std::sin(e);
            if (dot > best_dot) { best_dot = dot; best = i; }
        }
        return best;
    }

    // (left_ir, right_ir) for a direction — MySofa::get_hrtf_irs (src/sofa/loader.rs:136-199)
    std::pair<std::vector<float>, std::vector<float>> get_hrtf_irs(float azimuth_deg, float elevation_deg, float /*radius_m*/ = 1.0f) const {
        if (n_ == 0) throw SofaError("Filter length is zero.");
        const size_t i = nearest(azimuth_deg, elevation_deg);
        const float* l = ir_.data() + (i * 2) * n_;
        return {std::vector<float>(l, l + n_), std::vector<float>(l + n_, l + 2 * n_)};
    }

  private:
    std::vector<float> ir_, pos_;
    size_t m_ = 0, n_ = 0;
    float source_samplerate_ = 0.f, resampled_samplerate_ = 0.f;
};

// the plugin's speaker azimuth (negative = left, src/lib.rs:429-431) -> SOFA azimuth (positive = left)
inline float ui_azimuth_to_sofa(float ui_azimuth_deg) {
    float a = std::fmod(-ui_azimuth_deg, 360.0f);
    return a < 0 ? a + 360.0f : a;
}

// Left speaker direction -> (LSL, LSR), right speaker direction -> (RSL, RSR): four set_ir calls on anything with the
// reference's `set_ir(ConvolutionPath, ir)` (dsp.hpp ohs::ConvolutionEngine).  Returns the two measurement indices.
template <class Engine, class Path>
std::pair<size_t, size_t> wire_speakers(Engine& engine, const MySofa& hrirs, float az_left_deg, float el_left_deg, float az_right_deg,
                                        float el_right_deg, Path lsl, Path lsr, Path rsl, Path rsr) {
    const size_t il = hrirs.nearest(az_left_deg, el_left_deg), ir = hrirs.nearest(az_right_deg, el_right_deg);
    const auto left = hrirs.get_hrtf_irs(az_left_deg, el_left_deg), right = hrirs.get_hrtf_irs(az_right_deg, el_right_deg);
    engine.set_ir(lsl, left.first);
    engine.set_ir(lsr, left.second);
    engine.set_ir(rsl, right.first);
    engine.set_ir(rsr, right.second);
    return {il, ir};
}

}  // namespace ohs
