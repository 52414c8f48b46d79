"""ctypes binding of oracle/_build/liboracle.so (the C++ restatement in same_oracle.hpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "..", "tests", "golden")
_LIB = None

EV_NAMES = {
    0: "link.NoCarrier", 1: "link.Searching", 2: "link.Reading", 3: "link.Burst",
    16: "transport.Idle", 17: "transport.Assembling", 18: "transport.Message.SOM", 19: "transport.Message.EOM",
    20: "transport.Message.Err",
}


class OracleConfig(C.Structure):
    """Field-for-field the same as include/same_engine.h:same_config (mirrors builder.rs:50-67, 369-376)."""

    _fields_ = [
        ("input_rate", C.c_uint32),
        ("dc_blocker_len", C.c_float),
        ("agc_bandwidth", C.c_float),
        ("agc_gain_min", C.c_float),
        ("agc_gain_max", C.c_float),
        ("timing_bw_unlocked", C.c_float),
        ("timing_bw_locked", C.c_float),
        ("timing_max_deviation", C.c_float),
        ("squelch_power_open", C.c_float),
        ("squelch_power_close", C.c_float),
        ("squelch_bandwidth", C.c_float),
        ("preamble_max_errors", C.c_uint32),
        ("eq_enabled", C.c_uint32),
        ("eq_nff", C.c_uint32),
        ("eq_nfb", C.c_uint32),
        ("eq_relaxation", C.c_float),
        ("eq_regularization", C.c_float),
        ("frame_prefix_max_errors", C.c_uint32),
        ("frame_max_invalid_bytes", C.c_uint32),
    ]


class _Event(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("err", C.c_uint32), ("input_sample_counter", C.c_uint64),
        ("symbol_count", C.c_uint64), ("data_len", C.c_uint32), ("parity_errors", C.c_uint32),
        ("voting_bytes", C.c_uint32), ("reserved", C.c_uint32),
    ]


class _Soft(C.Structure):
    _fields_ = [("sample", C.c_uint64), ("zero", C.c_float), ("sym", C.c_float)]


class _Derived(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "sps", "agc_bw", "agc_gain0", "samples_per_ted", "period_min", "period_max",
        "alpha_unlocked", "beta_unlocked", "alpha_locked", "beta_locked")] + [("dc_len", C.c_uint32), ("ntaps", C.c_uint32)]


class OracleEvent:
    __slots__ = ("kind", "err", "sample", "symbol_count", "data", "parity_errors", "voting_bytes")

    def __init__(self, kind, err, sample, symbol_count, data, parity_errors, voting_bytes):
        self.kind, self.err, self.sample, self.symbol_count = kind, err, sample, symbol_count
        self.data, self.parity_errors, self.voting_bytes = data, parity_errors, voting_bytes

    @property
    def name(self):
        return EV_NAMES.get(self.kind, str(self.kind))

    @property
    def is_message(self):
        return self.kind in (18, 19)

    def key(self):
        """Everything that must match bit-for-bit between oracle and engine."""
        return (self.kind, self.err, self.sample, self.symbol_count, bytes(self.data), self.parity_errors, self.voting_bytes)

    def to_json(self):
        return {"kind": self.kind, "err": self.err, "sample": self.sample, "symbol_count": self.symbol_count,
                "data": self.data.hex(), "parity_errors": self.parity_errors, "voting_bytes": self.voting_bytes}

    def __repr__(self):
        return f"<{self.name} @{self.sample} sym {self.symbol_count} {self.data[:48]!r}>"


def build_oracle(force=False):
    """Compile oracle/_build/* with oracle/Makefile (g++ only; no reference sources are involved)."""
    so = os.path.join(HERE, "_build", "liboracle.so")
    if force or not os.path.exists(so) or any(
        os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(so)
        for f in ("same_oracle.hpp", "same_oracle_capi.cpp")
    ):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(build_oracle())
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.POINTER(OracleConfig)]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_reset.argtypes = [C.c_void_p]
        lib.oracle_enable_trace.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_process_s16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_process_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_process_zeros.argtypes = [C.c_void_p, C.c_size_t]
        lib.oracle_flush_samedec.argtypes = [C.c_void_p]
        lib.oracle_input_sample_counter.restype = C.c_uint64
        lib.oracle_input_sample_counter.argtypes = [C.c_void_p]
        lib.oracle_num_events.restype = C.c_size_t
        lib.oracle_num_events.argtypes = [C.c_void_p]
        lib.oracle_get_event.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_Event), C.c_void_p, C.c_size_t]
        lib.oracle_trace_len.restype = C.c_size_t
        lib.oracle_trace_len.argtypes = [C.c_void_p]
        lib.oracle_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_default_config.argtypes = [C.POINTER(OracleConfig), C.c_uint32, C.c_int]
        lib.oracle_get_derived.argtypes = [C.POINTER(OracleConfig), C.POINTER(_Derived), C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_decode_batch.restype = C.c_double
        lib.oracle_decode_batch.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                            C.c_int, C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


def default_config(rate=22050, samedec=True):
    cfg = OracleConfig()
    _lib().oracle_default_config(C.byref(cfg), rate, 1 if samedec else 0)
    return cfg


class Oracle:
    """One CPU receiver (one stream).  `Oracle.samedec()` = crates/samedec/src/main.rs:29-37 configuration."""

    def __init__(self, cfg=None):
        self.cfg = cfg if cfg is not None else default_config()
        self._h = _lib().oracle_create(C.byref(self.cfg))
        if not self._h:
            raise ValueError("invalid receiver configuration")

    @classmethod
    def samedec(cls, rate=22050):
        return cls(default_config(rate, True))

    @classmethod
    def library_default(cls, rate=22050):
        return cls(default_config(rate, False))

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().oracle_destroy(self._h)
            self._h = None

    def reset(self):
        _lib().oracle_reset(self._h)

    def enable_trace(self, on=True):
        _lib().oracle_enable_trace(self._h, 1 if on else 0)

    def process_s16(self, samples):
        a = np.ascontiguousarray(samples, dtype=np.int16)
        _lib().oracle_process_s16(self._h, a.ctypes.data, a.size)

    def process_f32(self, samples):
        a = np.ascontiguousarray(samples, dtype=np.float32)
        _lib().oracle_process_f32(self._h, a.ctypes.data, a.size)

    def process_zeros(self, n):
        _lib().oracle_process_zeros(self._h, n)

    def flush_samedec(self):
        _lib().oracle_flush_samedec(self._h)

    @property
    def input_sample_counter(self):
        return _lib().oracle_input_sample_counter(self._h)

    def events(self):
        lib, out = _lib(), []
        ev, buf = _Event(), (C.c_uint8 * 4096)()
        for i in range(lib.oracle_num_events(self._h)):
            lib.oracle_get_event(self._h, i, C.byref(ev), buf, 4096)
            out.append(OracleEvent(ev.kind, ev.err, ev.input_sample_counter, ev.symbol_count,
                                   bytes(buf[: min(ev.data_len, 4096)]), ev.parity_errors, ev.voting_bytes))
        return out

    def messages(self):
        return [e.data.decode("ascii") for e in self.events() if e.is_message]

    def soft_trace(self):
        n = _lib().oracle_trace_len(self._h)
        arr = (_Soft * n)()
        _lib().oracle_get_trace(self._h, arr, n)
        a = np.frombuffer(arr, dtype=np.dtype([("sample", "<u8"), ("zero", "<f4"), ("sym", "<f4")]))
        return a.copy()

    @staticmethod
    def derived(cfg):
        d = _Derived()
        mark = np.zeros(2 * 256, np.float32)
        space = np.zeros(2 * 256, np.float32)
        _lib().oracle_get_derived(C.byref(cfg), C.byref(d), mark.ctypes.data, space.ctypes.data, 256)
        out = {n: getattr(d, n) for n, _ in _Derived._fields_}
        out["mark"] = mark[: 2 * d.ntaps].reshape(-1, 2).copy()
        out["space"] = space[: 2 * d.ntaps].reshape(-1, 2).copy()
        return out

    @staticmethod
    def decode_batch(cfg, samples_2d, n_threads):
        """CPU baseline: one receiver per row of `samples_2d` (int16 [n_streams, len]); returns (seconds, bursts, msgs)."""
        a = np.ascontiguousarray(samples_2d, dtype=np.int16)
        nb = np.zeros(a.shape[0], np.uint32)
        nm = np.zeros(a.shape[0], np.uint32)
        secs = _lib().oracle_decode_batch(C.byref(cfg), a.ctypes.data, a.shape[0], a.shape[1], a.shape[1],
                                          int(n_threads), nb.ctypes.data, nm.ctypes.data)
        return secs, nb, nm


def load_golden_recording(name):
    """int16 samples of one of the reference's sample/ recordings (tests/golden/<name>.22050.s16le.bin.gz)."""
    with gzip.open(os.path.join(GOLDEN_DIR, f"{name}.22050.s16le.bin.gz"), "rb") as f:
        return np.frombuffer(f.read(), dtype="<i2").copy()
