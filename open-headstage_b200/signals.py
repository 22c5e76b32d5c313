"""Seeded synthetic inputs for the BASELINE.json configs: pink-noise streams, synthetic HRIR/BRIR
sets and the "typical" AutoEQ-like 10-band preset.  Pure numpy; used by bench.py, smoke() and the tests so that
the GPU path, the oracle and the CPU baseline all see identical bits."""
from __future__ import annotations

import numpy as np

# FilterType order of the reference (src/dsp/parametric_eq.rs:25-35)
PEAK, LOWSHELF, HIGHSHELF, LOWPASS, HIGHPASS, BANDPASS, NOTCH, ALLPASS = range(8)

# (type, fc Hz, Q, gain dB) — AutoEQ-like preset, every band enabled
EQ_PRESET_TYPICAL = [
    (LOWSHELF, 105.0, 0.7, 6.5),
    (PEAK, 60.0, 1.2, -3.0),
    (PEAK, 200.0, 0.9, -4.1),
    (PEAK, 800.0, 1.4, 2.2),
    (PEAK, 1800.0, 2.0, -2.5),
    (PEAK, 3200.0, 3.0, 4.0),
    (PEAK, 5500.0, 4.0, -5.0),
    (PEAK, 7400.0, 5.0, 3.1),
    (PEAK, 9800.0, 2.5, -2.0),
    (HIGHSHELF, 10000.0, 0.7, -4.0),
]

# Q=10 peaks at the bottom of the band, +-16 dB: the worst case for f32 DF2T round-off
EQ_PRESET_HARSH = [
    (PEAK, 20.0, 10.0, 16.0), (PEAK, 25.0, 10.0, -16.0), (PEAK, 30.0, 10.0, 16.0), (PEAK, 35.0, 10.0, -16.0),
    (PEAK, 40.0, 10.0, 16.0), (LOWSHELF, 50.0, 0.7, 12.0), (HIGHSHELF, 8000.0, 0.7, -12.0), (NOTCH, 1000.0, 8.0, 0.0),
    (ALLPASS, 500.0, 0.7, 0.0), (PEAK, 12000.0, 6.0, 9.0),
]


def pink_noise(n_frames: int, seed: int) -> np.ndarray:
    """White N(0,1) from default_rng(seed), shaped 1/sqrt(f) in the rFFT domain, peak-normalised to 1.0 (full scale)."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(n_frames)
    spec = np.fft.rfft(w)
    f = np.arange(spec.size, dtype=np.float64)
    f[0] = 1.0
    spec /= np.sqrt(f)
    spec[0] = 0.0
    x = np.fft.irfft(spec, n=n_frames)
    x /= np.max(np.abs(x))
    return x.astype(np.float32)


def stream_inputs(n_streams: int, n_frames: int, base_seed: int = 1000, unique: int | None = None) -> np.ndarray:
    """x[stream, channel, frame]; seed = base_seed + 2*stream + channel.  `unique` < n_streams tiles that many
    distinct streams (bench-only shortcut to keep host-side generation short; parity tests use unique streams)."""
    u = n_streams if unique is None else min(unique, n_streams)
    base = np.empty((u, 2, n_frames), np.float32)
    for s in range(u):
        for c in range(2):
            base[s, c] = pink_noise(n_frames, base_seed + 2 * s + c)
    if u == n_streams:
        return base
    reps = (n_streams + u - 1) // u
    return np.ascontiguousarray(np.tile(base, (reps, 1, 1))[:n_streams])


def synthetic_hrir_set(taps: int, decay: float, seed: int = 7) -> np.ndarray:
    """[4, taps] (LSL, LSR, RSL, RSR): default_rng(seed) N(0,1) * exp(-n/decay), each path unit-L2."""
    rng = np.random.default_rng(seed)
    n = np.arange(taps, dtype=np.float64)
    h = rng.standard_normal((4, taps)) * np.exp(-n / decay)
    h /= np.sqrt(np.sum(h * h, axis=1, keepdims=True))
    return h.astype(np.float32)


# BASELINE.json configs -> engine parameters
CONFIGS = {
    1: dict(name="cfg1: 1 stereo stream, 48 kHz, SOFA 200-tap HRIR, block 512", n_streams=1, fs=48000.0, block=512, taps=200),
    2: dict(name="cfg2: 1024 stereo streams, 48 kHz, 256-tap HRIR, block 256, 10-band PEQ", n_streams=1024, fs=48000.0,
            block=256, taps=256, decay=40.0),
    3: dict(name="cfg3: 65536 stereo streams (8192/GPU), 48 kHz, 512-tap HRIR, block 128", n_streams=8192, fs=48000.0,
            block=128, taps=512, decay=80.0),
    5: dict(name="cfg5: 2048 stereo streams (256/GPU), 96 kHz, 48000-tap BRIR, partition 1024", n_streams=256, fs=96000.0,
            block=1024, taps=48000, decay=0.15 * 96000.0),
}
