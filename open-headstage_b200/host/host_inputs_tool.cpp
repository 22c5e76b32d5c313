// host_inputs_tool — command-line face of the C++ host-side inputs (sofa.hpp, autoeq.hpp), used by the tests:
//   host_inputs_tool sofa <file> <M> <N> <az_left> <el_left> <az_right> <el_right>
//       reads the SOFA file without libmysofa/HDF5, prints M, N, the two selected indices and checksums (no GPU)
//   host_inputs_tool autoeq <csv>
//       prints the parsed bands (no GPU)
//   host_inputs_tool render <file> <M> <N> <csv> <in.f32> <out.f32>
//       GPU: SOFA (30 deg / 330 deg speakers) -> four set_ir calls, AutoEQ CSV -> update_band_coeffs, then the chain
//       EQ -> convolution over the [2][n_frames] f32 input file through the C++ mirror objects; writes [2][n_frames] f32
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "autoeq.hpp"
#include "dsp.hpp"
#include "sofa.hpp"

using namespace ohs;

static double checksum(const std::vector<float>& v) {
    double s = 0.0;
    for (size_t i = 0; i < v.size(); ++i) s += (double)v[i] * (double)((i % 7) + 1);
    return s;
}

int main(int argc, char** argv) {
    try {
        if (argc >= 9 && !std::strcmp(argv[1], "sofa")) {
            MySofa s = MySofa::open(argv[2], 48000.0f, (size_t)std::atol(argv[3]), (size_t)std::atol(argv[4]));
            const float azl = (float)std::atof(argv[5]), ell = (float)std::atof(argv[6]), azr = (float)std::atof(argv[7]), elr = (float)std::atof(argv[8]);
            const auto l = s.get_hrtf_irs(azl, ell), r = s.get_hrtf_irs(azr, elr);
            std::printf("{\"M\": %zu, \"N\": %zu, \"left_index\": %zu, \"right_index\": %zu, \"ir_checksum\": %.17g, \"pos_checksum\": %.17g, "
                        "\"left_l_checksum\": %.17g, \"right_r_checksum\": %.17g, \"ui_minus30\": %.9g, \"ui_plus30\": %.9g}\n",
                        s.measurements(), s.filter_length(), s.nearest(azl, ell), s.nearest(azr, elr), checksum(s.ir()), checksum(s.position()),
                        checksum(l.first), checksum(r.second), ui_azimuth_to_sofa(-30.0f), ui_azimuth_to_sofa(30.0f));
            return 0;
        }
        if (argc >= 3 && !std::strcmp(argv[1], "autoeq")) {
            const auto bands = parse_autoeq_csv(argv[2]);
            std::printf("[");
            for (size_t i = 0; i < bands.size(); ++i)
                std::printf("%s{\"enabled\": %s, \"filter_type\": %d, \"frequency\": %.9g, \"q\": %.9g, \"gain\": %.9g}", i ? ", " : "",
                            bands[i].enabled ? "true" : "false", (int)bands[i].filter_type, bands[i].frequency, bands[i].q, bands[i].gain);
            std::printf("]\n");
            return 0;
        }
        if (argc >= 8 && !std::strcmp(argv[1], "render")) {
            MySofa s = MySofa::open(argv[2], 48000.0f, (size_t)std::atol(argv[3]), (size_t)std::atol(argv[4]));
            const auto bands = parse_autoeq_csv(argv[5]);
            std::vector<float> l, r;
            {
                FILE* fi = std::fopen(argv[6], "rb");
                if (!fi) throw std::runtime_error("cannot read input");
                std::fseek(fi, 0, SEEK_END);
                const size_t total = (size_t)std::ftell(fi) / sizeof(float);
                std::fseek(fi, 0, SEEK_SET);
                l.resize(total / 2); r.resize(total / 2);
                if (std::fread(l.data(), sizeof(float), l.size(), fi) != l.size() || std::fread(r.data(), sizeof(float), r.size(), fi) != r.size())
                    throw std::runtime_error("short input file");
                std::fclose(fi);
            }
            const size_t n = l.size();
            ConvolutionEngine conv(512, 4096);
            const auto idx = wire_speakers(conv, s, ui_azimuth_to_sofa(-30.0f), 0.0f, ui_azimuth_to_sofa(30.0f), 0.0f, ConvolutionPath::Lsl,
                                           ConvolutionPath::Lsr, ConvolutionPath::Rsl, ConvolutionPath::Rsr);
            StereoParametricEQ eq(10, 48000.0f);
            apply_to_eq(eq, 48000.0f, bands);
            std::vector<float> ol(n), orr(n);
            eq.process_block(l, r);                 // Plugin::process order: EQ in place, then the convolver (src/lib.rs:1179-1200)
            conv.process_block(l, r, ol, orr);
            FILE* f = std::fopen(argv[7], "wb");
            if (!f) throw std::runtime_error("cannot write output");
            std::fwrite(ol.data(), sizeof(float), n, f);
            std::fwrite(orr.data(), sizeof(float), n, f);
            std::fclose(f);
            std::printf("{\"left_index\": %zu, \"right_index\": %zu, \"bands\": %zu}\n", idx.first, idx.second, bands.size());
            return 0;
        }
        std::fprintf(stderr, "usage: host_inputs_tool sofa|autoeq|render ... (see the source header)\n");
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
