#!/usr/bin/env python
"""Writes the INPUT side of the biquad golden vectors (committed): tests/golden/biquad_grid.txt (filter designs as f32
bit patterns: all 8 filter types x a seeded parameter grid, plus the "typical" and "harsh" 10-band cascades at 48 and
96 kHz) and tests/golden/biquad_input.f32 (8192 samples of the seeded pink noise, raw little-endian f32).
tools/gen_biquad_golden (Rust, real biquad 0.4.2) turns them into tests/golden/biquad_ref.txt."""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _bootstrap  # noqa: E402

S = _bootstrap.load_package().signals


def hx(v) -> str:
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(np.float32(v))))[0]


def main():
    lines = ["# record formats: tools/gen_biquad_golden/src/main.rs; numbers are f32 bit patterns (hex), decimal copies behind '#'"]
    rng = np.random.default_rng(20240607)
    for t in range(8):
        for _ in range(40):
            fs = float(rng.choice([44100.0, 48000.0, 96000.0]))
            fc = float(np.float32(rng.uniform(20.0, 0.49 * fs)))
            q = float(np.float32(rng.uniform(0.1, 10.0)))
            g = float(np.float32(rng.uniform(-16.0, 16.0)))
            lines.append("design %d %s %s %s %s   # fs %g fc %g q %g gain %g" % (t, hx(fs), hx(fc), hx(q), hx(g), fs, fc, q, g))
    # error paths of from_params: above Nyquist, negative Q
    lines.append("design 0 %s %s %s %s   # OutsideNyquist" % (hx(48000.0), hx(30000.0), hx(1.0), hx(0.0)))
    lines.append("design 0 %s %s %s %s   # NegativeQ" % (hx(48000.0), hx(1000.0), hx(-0.5), hx(0.0)))
    n = 8192
    for name, preset in (("typical", S.EQ_PRESET_TYPICAL), ("harsh", S.EQ_PRESET_HARSH)):
        for fs in (48000.0, 96000.0):
            lines.append("cascade %s_%d %s %d %d" % (name, int(fs), hx(fs), len(preset), n))
            for (t, fc, q, g) in preset:
                lines.append("band %d %s %s %s   # fc %g q %g gain %g" % (t, hx(fc), hx(q), hx(g), fc, q, g))
    open(os.path.join(HERE, "biquad_grid.txt"), "w").write("\n".join(lines) + "\n")
    S.pink_noise(n, 1).astype("<f4").tofile(os.path.join(HERE, "biquad_input.f32"))
    print("wrote biquad_grid.txt (%d records) and biquad_input.f32 (%d samples)" % (len(lines) - 1, n))


if __name__ == "__main__":
    main()
