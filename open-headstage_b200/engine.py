"""ctypes binding of include/ohs.h plus Python mirrors of the reference's two DSP objects.

`Engine` is the batched handle (n_streams chains).  `ConvolutionEngine` and `StereoParametricEQ` keep the reference's
method names and argument meaning (src/dsp/convolution.rs:87,111,141; src/dsp/parametric_eq.rs:132,144,166,181) over
a one-stream engine, so parity tests read like the reference's own unit tests.  Everything computes on the GPU through
libohs_cuda.so; if the library or a CUDA device is missing these raise — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _build

OHS_ALL = -1
ABI_VERSION = 2   # OHS_ABI_VERSION of include/ohs.h this binding was written against
LSL, LSR, RSL, RSR = 0, 1, 2, 3
PEAK, LOWSHELF, HIGHSHELF, LOWPASS, HIGHPASS, BANDPASS, NOTCH, ALLPASS = range(8)

_f32p = C.POINTER(C.c_float)


class OhsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("ohs error %d: %s" % (code, msg))
        self.code = code


class _Config(C.Structure):
    _fields_ = [("n_streams", C.c_int32), ("block", C.c_int32), ("max_taps", C.c_int32), ("n_bands", C.c_int32),
                ("n_hrir_sets", C.c_int32), ("n_eq_sets", C.c_int32), ("device", C.c_int32), ("sample_rate", C.c_float)]


# every symbol include/ohs.h declares: (name, restype, argtypes)
_VP = C.c_void_p
SYMBOLS = [
    ("ohs_create", C.c_int, [C.POINTER(_Config), C.POINTER(_VP)]),
    ("ohs_destroy", C.c_int, [_VP]),
    ("ohs_abi_version", C.c_int, []),
    ("ohs_last_error", C.c_char_p, []),
    ("ohs_set_ir", C.c_int, [_VP, C.c_int, C.c_int, _f32p, C.c_size_t]),
    ("ohs_num_partitions", C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    ("ohs_bind_stream_hrir", C.c_int, [_VP, C.c_int, C.c_int]),
    ("ohs_commit_filters", C.c_int, [_VP]),
    ("ohs_filter_table", C.c_int, [_VP, C.POINTER(_VP), C.POINTER(C.c_size_t)]),
    ("ohs_mark_filters_external", C.c_int, [_VP, C.c_int, C.c_int]),
    ("ohs_eq_design", C.c_int, [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _f32p]),
    ("ohs_eq_update_band", C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    ("ohs_eq_set_band", C.c_int, [_VP, C.c_int, C.c_int, _f32p, C.c_int]),
    ("ohs_bind_stream_eq", C.c_int, [_VP, C.c_int, C.c_int]),
    ("ohs_eq_reset", C.c_int, [_VP]),
    ("ohs_eq_frequency_response", C.c_int, [_VP, C.c_int, C.c_float, _f32p, _f32p, C.c_size_t]),
    ("ohs_set_eq_enable", C.c_int, [_VP, C.c_int]),
    ("ohs_set_conv_enable", C.c_int, [_VP, C.c_int]),
    ("ohs_set_bypass", C.c_int, [_VP, C.c_int]),
    ("ohs_set_gain", C.c_int, [_VP, C.c_int, C.c_float]),
    ("ohs_conv_reset", C.c_int, [_VP]),
    ("ohs_set_time_batch", C.c_int, [_VP, C.c_int]),
    ("ohs_streams_per_cta", C.c_int, [_VP, C.POINTER(C.c_int)]),
    ("ohs_prepare", C.c_int, [_VP, C.c_size_t, C.c_int]),
    ("ohs_debug_trace", C.c_int, [_VP, _VP]),
    ("ohs_process_device", C.c_int, [_VP, _VP, _VP, C.c_size_t, C.c_size_t]),
    ("ohs_process", C.c_int, [_VP, _VP, _VP, C.c_size_t, C.c_size_t]),
    ("ohs_process_fifo", C.c_int, [_VP, _VP, _VP, C.c_size_t, C.c_size_t]),
    ("ohs_sync", C.c_int, [_VP]),
    ("ohs_cuda_stream", C.c_int, [_VP, C.POINTER(_VP)]),
    ("ohs_launch_count", C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    ("ohs_enable_timing", C.c_int, [_VP, C.c_int]),
    ("ohs_last_kernel_ms", C.c_int, [_VP, C.POINTER(C.c_float)]),
    ("ohs_mix_device", C.c_int, [_VP, _VP, _VP, C.c_size_t, C.c_size_t, C.c_size_t]),
    ("ohs_comm_unique_id", C.c_int, [_VP]),
    ("ohs_comm_create", C.c_int, [C.POINTER(_VP), C.c_int, C.c_int, C.c_int, _VP]),
    ("ohs_comm_destroy", C.c_int, [_VP]),
    ("ohs_broadcast_hrir", C.c_int, [_VP, _VP, C.c_int]),
    ("ohs_reduce_bus", C.c_int, [_VP, _VP, _VP, C.c_size_t, C.c_int]),
    ("ohs_host_alloc", C.c_int, [C.POINTER(_VP), C.c_size_t]),
    ("ohs_host_free", C.c_int, [_VP]),
    ("ohs_state_bytes", C.c_int, [_VP, C.POINTER(C.c_size_t)]),
    ("ohs_state_export", C.c_int, [_VP, _VP, C.c_size_t]),
    ("ohs_state_import", C.c_int, [_VP, _VP, C.c_size_t]),
]

_lib = None


def load_library(build: bool = True):
    """dlopen libohs_cuda.so (building it in-tree first if it is missing or stale and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build and _build.is_stale():
        try:
            _build.build_library()
        except Exception as e:
            if not os.path.exists(path):
                raise
            import warnings

            warnings.warn("libohs_cuda.so is older than its sources and could not be rebuilt (%s); loading the stale library" % e,
                          RuntimeWarning, stacklevel=2)
    if not os.path.exists(path):
        raise OhsError(-3, "libohs_cuda.so is missing and could not be built; there is no CPU fallback")
    L = C.CDLL(path)
    L.ohs_abi_version.restype = C.c_int
    got = L.ohs_abi_version()
    if got != ABI_VERSION:
        raise OhsError(-1, "libohs_cuda.so has ABI version %d, this binding expects %d (rebuild the library)" % (got, ABI_VERSION))
    for name, res, args in SYMBOLS:
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise OhsError(rc, load_library().ohs_last_error().decode())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def eq_design(filter_type: int, fs: float, fc: float, q: float, gain_db: float) -> np.ndarray:
    """Host-side RBJ design identical to biquad 0.4.2 from_params (src/dsp/parametric_eq.rs:105-111)."""
    out = np.zeros(5, np.float32)
    _check(load_library().ohs_eq_design(filter_type, fs, fc, q, gain_db, out.ctypes.data_as(_f32p)))
    return out


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (one rank calls it and ships the 128 bytes to the others)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(load_library().ohs_comm_unique_id(buf))
    return buf.raw


class Comm:
    """ohs_comm: an NCCL communicator owned by the C ABI library (ohs_broadcast_hrir / ohs_reduce_bus)."""

    def __init__(self, world: int, rank: int, device: int, unique_id: bytes):
        assert len(unique_id) == COMM_ID_BYTES
        self._L = load_library()
        self.world, self.rank, self.device = world, rank, device
        self._c = _VP()
        _check(self._L.ohs_comm_create(C.byref(self._c), world, rank, device, unique_id))

    def close(self):
        if getattr(self, "_c", None):
            self._L.ohs_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """cudaHostAlloc'd float32 array (ohs_host_alloc) viewed as numpy."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape))
        self._ptr = _VP()
        _check(load_library().ohs_host_alloc(C.byref(self._ptr), n * 4))
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, _f32p), shape=(n,)).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            load_library().ohs_host_free(self._ptr)
            self._ptr = _VP()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One ohs_engine handle: n_streams independent EQ -> 4-path convolution -> gain chains resident on one GPU."""

    def __init__(self, n_streams: int, block: int, max_taps: int, n_bands: int = 10, n_hrir_sets: int = 1,
                 n_eq_sets: int = 1, device: int = 0, sample_rate: float = 48000.0):
        self._L = load_library()
        self.n_streams, self.block, self.max_taps, self.n_bands = n_streams, block, max_taps, n_bands
        self.n_hrir_sets, self.n_eq_sets, self.device = n_hrir_sets, n_eq_sets, device
        self.sample_rate = sample_rate
        cfg = _Config(n_streams, block, max_taps, n_bands, n_hrir_sets, n_eq_sets, device, sample_rate)
        self._h = _VP()
        _check(self._L.ohs_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.ohs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- HRIR
    def set_ir(self, path: int, ir, hrir_set: int = 0):
        ir = _f32(ir).ravel()
        _check(self._L.ohs_set_ir(self._h, hrir_set, path, ir.ctypes.data_as(_f32p) if ir.size else None, ir.size))

    def set_hrir_set(self, irs, hrir_set: int = 0):
        for p in range(4):
            self.set_ir(p, irs[p], hrir_set)

    def num_partitions(self, path: int, hrir_set: int = 0) -> int:
        out = C.c_int()
        _check(self._L.ohs_num_partitions(self._h, hrir_set, path, C.byref(out)))
        return out.value

    def bind_stream_hrir(self, stream: int, hrir_set: int):
        _check(self._L.ohs_bind_stream_hrir(self._h, stream, hrir_set))

    def commit_filters(self):
        _check(self._L.ohs_commit_filters(self._h))

    def filter_table(self):
        p, n = _VP(), C.c_size_t()
        _check(self._L.ohs_filter_table(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def mark_filters_external(self, hrir_set: int, partitions: int):
        _check(self._L.ohs_mark_filters_external(self._h, hrir_set, partitions))

    def broadcast_hrir(self, comm: "Comm", root: int = 0):
        """One HRIR spectra table for the whole job: rank `root` has called set_ir, everybody receives (C ABI, NCCL)."""
        _check(self._L.ohs_broadcast_hrir(self._h, comm._c, root))

    def reduce_bus(self, comm: "Comm", d_bus: int, n_floats: int, root: int = 0):
        """In-place NCCL sum of the per-GPU buses onto `root`, on the engine's stream (config 4)."""
        _check(self._L.ohs_reduce_bus(self._h, comm._c, d_bus, n_floats, root))

    # ---- EQ
    def eq_update_band(self, band: int, filter_type: int, fc: float, q: float, gain_db: float, enabled: bool = True, eq_set: int = 0):
        _check(self._L.ohs_eq_update_band(self._h, eq_set, band, filter_type, fc, q, gain_db, int(enabled)))

    def eq_set_band(self, band: int, coeffs, enabled: bool = True, eq_set: int = 0):
        c = _f32(coeffs)
        _check(self._L.ohs_eq_set_band(self._h, eq_set, band, c.ctypes.data_as(_f32p), int(enabled)))

    def eq_set_preset(self, preset, eq_set: int = 0, enabled: bool = True):
        for i, (t, fc, q, g) in enumerate(preset):
            self.eq_update_band(i, t, fc, q, g, enabled, eq_set)

    def bind_stream_eq(self, stream: int, eq_set: int):
        _check(self._L.ohs_bind_stream_eq(self._h, stream, eq_set))

    def eq_reset(self):
        _check(self._L.ohs_eq_reset(self._h))

    def eq_frequency_response(self, freqs, eq_set: int = 0, sample_rate: float = 0.0) -> np.ndarray:
        f = _f32(freqs)
        out = np.zeros_like(f)
        _check(self._L.ohs_eq_frequency_response(self._h, eq_set, sample_rate, f.ctypes.data_as(_f32p), out.ctypes.data_as(_f32p), f.size))
        return out

    # ---- switches
    def set_eq_enable(self, on: bool):
        _check(self._L.ohs_set_eq_enable(self._h, int(on)))

    def set_conv_enable(self, on: bool):
        _check(self._L.ohs_set_conv_enable(self._h, int(on)))

    def set_bypass(self, on: bool):
        _check(self._L.ohs_set_bypass(self._h, int(on)))

    def set_gain(self, gain: float, stream: int = OHS_ALL):
        _check(self._L.ohs_set_gain(self._h, stream, gain))

    def conv_reset(self):
        _check(self._L.ohs_conv_reset(self._h))

    def set_time_batch(self, on: bool):
        _check(self._L.ohs_set_time_batch(self._h, int(on)))

    def streams_per_cta(self) -> int:
        out = C.c_int()
        _check(self._L.ohs_streams_per_cta(self._h, C.byref(out)))
        return out.value

    def prepare(self, n_frames: int, host_io: bool = False):
        """Upload pending set-up and allocate scratch/staging for calls of n_frames, outside any timed region."""
        _check(self._L.ohs_prepare(self._h, n_frames, int(host_io)))

    def debug_trace(self, d_stamps: int):
        _check(self._L.ohs_debug_trace(self._h, d_stamps))

    # ---- processing
    def process(self, x, out=None) -> np.ndarray:
        """Host flavour.  x[stream, channel, frame] float32; returns the rendered array (or fills `out`)."""
        x = _f32(x)
        assert x.ndim == 3 and x.shape[0] == self.n_streams and x.shape[1] == 2, x.shape
        if out is None:
            y = np.empty_like(x)
        else:
            y = out
            if not (isinstance(y, np.ndarray) and y.dtype == np.float32 and y.shape == x.shape and y.flags["C_CONTIGUOUS"]
                    and y.flags["WRITEABLE"]):
                raise ValueError("out must be a writeable C-contiguous float32 array of shape %s" % (x.shape,))
        _check(self._L.ohs_process(self._h, x.ctypes.data, y.ctypes.data, x.shape[2], x.shape[2]))
        return y

    def process_fifo(self, x) -> np.ndarray:
        x = _f32(x)
        assert x.ndim == 3 and x.shape[0] == self.n_streams and x.shape[1] == 2, x.shape
        y = np.empty_like(x)
        _check(self._L.ohs_process_fifo(self._h, x.ctypes.data, y.ctypes.data, x.shape[2], x.shape[2]))
        return y

    def process_device(self, d_in: int, d_out: int, n_frames: int, row_stride: int | None = None):
        """Device flavour: raw device pointers (e.g. torch.Tensor.data_ptr()); enqueues on the engine's stream."""
        _check(self._L.ohs_process_device(self._h, d_in, d_out, n_frames, n_frames if row_stride is None else row_stride))

    def mix_device(self, d_in: int, d_bus: int, n_frames: int, row_stride: int | None = None, bus_stride: int | None = None):
        """Sum this engine's rendered streams into a stereo bus [2][bus_stride] (config 4 object mixdown)."""
        _check(self._L.ohs_mix_device(self._h, d_in, d_bus, n_frames, n_frames if row_stride is None else row_stride,
                                      n_frames if bus_stride is None else bus_stride))

    def sync(self):
        _check(self._L.ohs_sync(self._h))

    def cuda_stream(self) -> int:
        s = _VP()
        _check(self._L.ohs_cuda_stream(self._h, C.byref(s)))
        return s.value or 0

    def launch_count(self) -> int:
        n = C.c_uint64()
        _check(self._L.ohs_launch_count(self._h, C.byref(n)))
        return n.value

    def enable_timing(self, on: bool = True):
        _check(self._L.ohs_enable_timing(self._h, int(on)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _check(self._L.ohs_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    # ---- state
    def state_export(self) -> bytes:
        n = C.c_size_t()
        _check(self._L.ohs_state_bytes(self._h, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(self._L.ohs_state_export(self._h, buf, n.value))
        return buf.raw

    def state_import(self, blob: bytes):
        _check(self._L.ohs_state_import(self._h, blob, len(blob)))


# ---------------------------------------------------------------------------------------------------------------
# Mirrors of the reference's objects (one stereo stream each)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class BandConfig:
    """src/dsp/parametric_eq.rs:37-44"""
    filter_type: int
    center_freq: float
    q: float
    gain_db: float
    enabled: bool


class ConvolutionEngine:
    """src/dsp/convolution.rs ConvolutionEngine::{new, set_ir, process_block}; block = BLOCK_SIZE (:22, 512 there)."""

    def __init__(self, block: int = 512, max_taps: int = 4096, device: int = 0):
        self._e = Engine(1, block, max_taps, n_bands=0, device=device)
        self.block = block

    def set_ir(self, path: int, ir_data):
        self._e.set_ir(path, ir_data)

    def num_partitions(self, path: int) -> int:
        return self._e.num_partitions(path)

    def process_block(self, input_left, input_right):
        """Any host-block length, with the reference's FIFO/zero-fill semantics (:141-182)."""
        x = np.stack([_f32(input_left), _f32(input_right)])[None]
        y = self._e.process_fifo(x)
        return y[0, 0], y[0, 1]


class StereoParametricEQ:
    """src/dsp/parametric_eq.rs StereoParametricEQ::{new, update_band_coeffs, process_block, reset_all_bands_state}."""

    def __init__(self, num_bands: int, initial_sample_rate: float, device: int = 0):
        self._e = Engine(1, 256, 1, n_bands=num_bands, device=device, sample_rate=initial_sample_rate)
        self._e.set_conv_enable(False)
        self._e.set_eq_enable(True)
        self.num_bands = num_bands

    def update_band_coeffs(self, band_idx: int, sample_rate: float, config: BandConfig):
        c = eq_design(config.filter_type, sample_rate, config.center_freq, config.q, config.gain_db)
        self._e.eq_set_band(band_idx, c, config.enabled)

    def process_block(self, input_left, input_right):
        """In place in the reference (:166); returns the filtered copies."""
        x = np.stack([_f32(input_left), _f32(input_right)])[None]
        y = self._e.process(x)
        return y[0, 0], y[0, 1]

    def reset_all_bands_state(self):
        self._e.eq_reset()

    def calculate_frequency_response(self, sample_rate: float, frequencies):
        return self._e.eq_frequency_response(frequencies, sample_rate=sample_rate)
