// autoeq.hpp — AutoEQ parametric-EQ CSV ingestion in C++: the host step in front of `update_band_coeffs` (SURVEY.md 8f
// rank 4).  Mirrors the reference's `parse_autoeq_csv` (src/autoeq_parser.rs:43-70): header `Filter-Type,Fc,Q,Gain`
// (columns found by name, any order), filter types PK / LS / HS -> Peak / LowShelf / HighShelf, anything else is an
// error ("Unsupported filter type: ..."); every parsed band is enabled.
#pragma once

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "dsp.hpp"

namespace ohs {

// src/autoeq_parser.rs:34-41
struct BandSetting {
    bool enabled = false;
    FilterType filter_type = FilterType::Peak;
    float frequency = 0.f, q = 0.f, gain = 0.f;
};

namespace detail {
inline std::string trim(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t\r\n\""), b = s.find_last_not_of(" \t\r\n\"");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline std::vector<std::string> split_csv(const std::string& line) {
    std::vector<std::string> out;
    std::string cell;
    std::istringstream ss(line);
    while (std::getline(ss, cell, ',')) out.push_back(trim(cell));
    if (!line.empty() && line.back() == ',') out.emplace_back();
    return out;
}
}  // namespace detail

// src/autoeq_parser.rs:43-50
inline FilterType map_filter_type(const std::string& autoeq_type) {
    if (autoeq_type == "PK") return FilterType::Peak;
    if (autoeq_type == "LS") return FilterType::LowShelf;
    if (autoeq_type == "HS") return FilterType::HighShelf;
    throw std::runtime_error("Unsupported filter type: " + autoeq_type);
}

inline std::vector<BandSetting> parse_autoeq_csv_text(const std::string& text) {
    std::istringstream in(text);
    std::string line;
    if (!std::getline(in, line)) throw std::runtime_error("empty AutoEQ CSV");
    const std::vector<std::string> head = detail::split_csv(line);
    int c_type = -1, c_fc = -1, c_q = -1, c_gain = -1;
    for (size_t i = 0; i < head.size(); ++i) {
        if (head[i] == "Filter-Type") c_type = (int)i;
        else if (head[i] == "Fc") c_fc = (int)i;
        else if (head[i] == "Q") c_q = (int)i;
        else if (head[i] == "Gain") c_gain = (int)i;
    }
    if (c_type < 0 || c_fc < 0 || c_q < 0 || c_gain < 0) throw std::runtime_error("AutoEQ CSV needs the columns Filter-Type,Fc,Q,Gain");
    std::vector<BandSetting> bands;
    while (std::getline(in, line)) {
        if (detail::trim(line).empty()) continue;
        const std::vector<std::string> cell = detail::split_csv(line);
        const size_t need = (size_t)std::max(std::max(c_type, c_fc), std::max(c_q, c_gain)) + 1;
        if (cell.size() < need) throw std::runtime_error("AutoEQ CSV row with too few columns: " + line);
        BandSetting b;
        b.enabled = true;
        b.filter_type = map_filter_type(cell[c_type]);
        b.frequency = std::stof(cell[c_fc]);
        b.q = std::stof(cell[c_q]);
        b.gain = std::stof(cell[c_gain]);
        bands.push_back(b);
    }
    return bands;
}

// parse_autoeq_csv(path) (src/autoeq_parser.rs:52-70)
inline std::vector<BandSetting> parse_autoeq_csv(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_autoeq_csv_text(ss.str());
}

// update_band_coeffs for each parsed band; bands beyond the equaliser's band count are ignored, as in the reference
// (src/dsp/parametric_eq.rs:145)
inline void apply_to_eq(StereoParametricEQ& eq, float sample_rate, const std::vector<BandSetting>& bands) {
    for (size_t i = 0; i < bands.size(); ++i) {
        BandConfig c;
        c.filter_type = bands[i].filter_type; c.center_freq = bands[i].frequency; c.q = bands[i].q; c.gain_db = bands[i].gain;
        c.enabled = bands[i].enabled;
        eq.update_band_coeffs(i, sample_rate, c);
    }
}

}  // namespace ohs
