"""Host-side HRIR source for the engine: the step immediately before `set_ir`.

The reference does this on the host in Rust over libmysofa (`MySofa::open` / `get_hrtf_irs`,
src/sofa/loader.rs:79,136) and — as shipped — never connects it to the convolver (SURVEY finding 5).  This module is
the host-side equivalent the batch renderer needs: it reads a SimpleFreeFieldHRIR SOFA file without an HDF5 library
(the two zlib-compressed, byte-shuffled datasets `SourcePosition` and `Data.IR` are located by scanning for zlib
streams of the right decoded size), picks the measurement nearest to a direction, and wires a stereo speaker pair
to the four convolution paths.  Pure numpy; no sample arithmetic happens here.

Conventions: SOFA spherical coordinates, azimuth in degrees counter-clockwise (positive = left), elevation in
degrees, radius in metres.  The plugin's UI uses negative azimuth = left (src/lib.rs:429-431), hence
`ui_azimuth_to_sofa`.  Selection is nearest neighbour on the unit sphere; libmysofa additionally interpolates between
neighbours and resamples to the processing rate — not reproduced (parity for selection is "unpinned"):
taps are used as measured.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass

import numpy as np


def _find_zlib_stream(buf: bytes, decoded_size: int) -> bytes:
    i = 0
    while True:
        i = buf.find(b"\x78", i)
        if i < 0 or i + 1 >= len(buf):
            raise ValueError("no zlib stream with decoded size %d" % decoded_size)
        # a zlib header is two bytes whose big-endian value is a multiple of 31 (RFC 1950): 78 01 / 78 5E / 78 9C / 78 DA
        if buf[i + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            try:
                # bounded: stop inflating as soon as the stream is longer than the dataset we are looking for
                out = zlib.decompressobj().decompress(buf[i:], decoded_size + 1)
                if len(out) == decoded_size:
                    return out
            except zlib.error:
                pass
        i += 1


def _unshuffle_f64(raw: bytes, shape) -> np.ndarray:
    return np.frombuffer(raw, np.uint8).reshape(8, -1).T.copy().view("<f8").reshape(shape)


@dataclass
class HrirSet:
    ir: np.ndarray          # [M, 2, N] float32: measurement, ear (0 = left, 1 = right), tap
    position: np.ndarray    # [M, 3] azimuth deg, elevation deg, radius m
    sample_rate: float

    @property
    def filter_length(self) -> int:
        return int(self.ir.shape[2])

    def nearest(self, azimuth_deg: float, elevation_deg: float) -> int:
        """Index of the measurement closest (great-circle) to the direction."""
        az, el = np.deg2rad(self.position[:, 0]), np.deg2rad(self.position[:, 1])
        a, e = np.deg2rad(azimuth_deg % 360.0), np.deg2rad(elevation_deg)
        dots = np.cos(el) * np.cos(e) * np.cos(az - a) + np.sin(el) * np.sin(e)
        return int(np.argmax(dots))

    def get_hrtf_irs(self, azimuth_deg: float, elevation_deg: float, radius_m: float = 1.0):
        """(left_ir, right_ir) for a direction — the shape of MySofa::get_hrtf_irs (src/sofa/loader.rs:136-199)."""
        i = self.nearest(azimuth_deg, elevation_deg)
        return self.ir[i, 0].copy(), self.ir[i, 1].copy()


def load_sofa(path: str, n_measurements: int | None = None, n_taps: int | None = None, sample_rate: float = 44100.0) -> HrirSet:
    """Read SourcePosition and Data.IR of a SimpleFreeFieldHRIR file with two receivers.  M and N are found by trying
    the common CIPIC/ARI/… shapes unless given."""
    buf = open(path, "rb").read()
    shapes = [(n_measurements, n_taps)] if n_measurements and n_taps else [(1250, 200), (1550, 256), (2304, 256), (828, 256), (710, 512), (2702, 512)]
    for m, n in shapes:
        try:
            pos = _unshuffle_f64(_find_zlib_stream(buf, m * 3 * 8), (m, 3))
            ir = _unshuffle_f64(_find_zlib_stream(buf, m * 2 * n * 8), (m, 2, n))
            return HrirSet(ir.astype(np.float32), pos.astype(np.float32), sample_rate)
        except ValueError:
            continue
    raise ValueError("could not locate SourcePosition / Data.IR datasets in %s" % path)


def from_arrays(ir, position, sample_rate: float) -> HrirSet:
    return HrirSet(np.asarray(ir, np.float32), np.asarray(position, np.float32), float(sample_rate))


def ui_azimuth_to_sofa(ui_azimuth_deg: float) -> float:
    """The plugin's speaker azimuth (negative = left, src/lib.rs:429-431) -> SOFA azimuth (positive = left)."""
    return (-ui_azimuth_deg) % 360.0


def wire_speakers(engine, hrirs: HrirSet, az_left_deg: float, el_left_deg: float, az_right_deg: float, el_right_deg: float,
                  hrir_set: int = 0):
    """The wiring the reference intends but never performs (github_issues/sofa_implement_logic_select_extract_hrirs.md):
    left speaker direction -> (LSL, LSR), right speaker direction -> (RSL, RSR), then four set_ir calls.
    Angles are SOFA azimuths/elevations.  Returns the two measurement indices used."""
    il = hrirs.nearest(az_left_deg, el_left_deg)
    ir_ = hrirs.nearest(az_right_deg, el_right_deg)
    engine.set_ir(0, hrirs.ir[il, 0], hrir_set)   # LSL
    engine.set_ir(1, hrirs.ir[il, 1], hrir_set)   # LSR
    engine.set_ir(2, hrirs.ir[ir_, 0], hrir_set)  # RSL
    engine.set_ir(3, hrirs.ir[ir_, 1], hrir_set)  # RSR
    return il, ir_
