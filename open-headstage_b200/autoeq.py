"""AutoEQ parametric-EQ CSV ingestion — the host step in front of `update_band_coeffs`.
Mirrors the reference's `parse_autoeq_csv` (src/autoeq_parser.rs:43-70): columns `Filter-Type,Fc,Q,Gain`, filter types
PK / LS / HS -> Peak / LowShelf / HighShelf, anything else is an error; every parsed band is enabled."""
from __future__ import annotations

import csv
import io
from dataclasses import dataclass

PEAK, LOWSHELF, HIGHSHELF = 0, 1, 2
_TYPES = {"PK": PEAK, "LS": LOWSHELF, "HS": HIGHSHELF}  # src/autoeq_parser.rs:43-50


@dataclass
class BandSetting:
    """src/autoeq_parser.rs:34-41"""
    enabled: bool
    filter_type: int
    frequency: float
    q: float
    gain: float


def parse_autoeq_csv(text_or_path: str) -> list[BandSetting]:
    if "\n" not in text_or_path and "," not in text_or_path:
        with open(text_or_path, newline="") as f:
            text = f.read()
    else:
        text = text_or_path
    bands = []
    for row in csv.DictReader(io.StringIO(text)):
        t = row["Filter-Type"].strip()
        if t not in _TYPES:
            raise ValueError("Unsupported filter type: %s" % t)
        bands.append(BandSetting(True, _TYPES[t], float(row["Fc"]), float(row["Q"]), float(row["Gain"])))
    return bands


def apply_to_engine(engine, bands, eq_set: int = 0):
    """update_band_coeffs for each parsed band (bands beyond the engine's n_bands are ignored, as in the reference)."""
    for i, b in enumerate(bands):
        engine.eq_update_band(i, b.filter_type, b.frequency, b.q, b.gain, b.enabled, eq_set)
