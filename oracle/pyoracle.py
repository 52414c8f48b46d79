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
        for f in ("same_oracle.hpp", "same_oracle_capi.cpp", "synth_cpu.hpp")
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
        lib.oracle_decode_batch_events.restype = C.c_void_p
        lib.oracle_decode_batch_events.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                   C.c_void_p, C.c_int, C.c_int]
        lib.oracle_batch_num_events.restype = C.c_size_t
        lib.oracle_batch_num_events.argtypes = [C.c_void_p]
        lib.oracle_batch_payload_bytes.restype = C.c_size_t
        lib.oracle_batch_payload_bytes.argtypes = [C.c_void_p]
        lib.oracle_batch_seconds.restype = C.c_double
        lib.oracle_batch_seconds.argtypes = [C.c_void_p]
        lib.oracle_batch_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_batch_free.argtypes = [C.c_void_p]
        lib.oracle_agc_block_model.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float,
                                               C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.oracle_synth_generate.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int]
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


# same layout as include/same_engine.h:same_event (and sameold_b200.SameBatchReceiver.EVENT_DTYPE)
BATCH_EVENT_DTYPE = np.dtype([("stream", "<u4"), ("seq", "<u4"), ("sample", "<u8"), ("symbol_count", "<u8"),
                              ("kind", "<u4"), ("err", "<u4"), ("data_offset", "<u4"), ("data_len", "<u4"),
                              ("parity_errors", "<u2"), ("voting_bytes", "<u2"), ("flags", "<u4")])


def decode_batch_events(cfg, samples_2d, n_threads, lengths=None, flush=False):
    """One oracle receiver per row of `samples_2d` (int16 [n_streams, stride]) on `n_threads` host threads; returns
    (events, payload, seconds): a structured array in the engine's same_event layout sorted by (stream, occurrence;
    seq counts from 0 per stream) and the concatenated payload bytes."""
    a = samples_2d
    assert a.dtype == np.int16 and a.ndim == 2 and a.strides[1] == 2
    stride = a.strides[0] // 2
    lens = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.uint32)
    h = _lib().oracle_decode_batch_events(C.byref(cfg), a.ctypes.data, a.shape[0], stride, a.shape[1],
                                          None if lens is None else lens.ctypes.data, int(n_threads), 1 if flush else 0)
    try:
        ev = np.zeros(_lib().oracle_batch_num_events(h), BATCH_EVENT_DTYPE)
        pay = np.zeros(max(_lib().oracle_batch_payload_bytes(h), 1), np.uint8)
        _lib().oracle_batch_copy(h, ev.ctypes.data, pay.ctypes.data)
        secs = _lib().oracle_batch_seconds(h)
        return ev, pay[: _lib().oracle_batch_payload_bytes(h)], secs
    finally:
        _lib().oracle_batch_free(h)


def agc_block_model(cfg, samples, block=2048, warm=1024, guess_gain=None):
    """(blocks, blocks whose warm-started gain differs bitwise from the sequential one, worst samples-to-coalesce):
    the CPU model of the long-stream path's speculative AGC (oracle_agc_block_model)."""
    a = np.ascontiguousarray(samples, dtype=np.int16)
    nb, bad, worst = C.c_uint32(), C.c_uint32(), C.c_uint32()
    g0 = min(1.0, cfg.agc_gain_min) if guess_gain is None else guess_gain
    _lib().oracle_agc_block_model(C.byref(cfg), a.ctypes.data, a.size, block, warm, C.c_float(g0), C.byref(nb), C.byref(bad),
                                  C.byref(worst))
    return nb.value, bad.value, worst.value


class _CpuBurst(C.Structure):
    _fields_ = [("start_sample", C.c_double), ("byte_offset", C.c_uint32), ("n_bytes", C.c_uint32)]


def synth_cpu(plans, n_samples, rate=22050, n_threads=1, amplitude=16384.0, noise_sigma=11585.0 / 10.0 ** 0.5):
    """CPU rendering (oracle/synth_cpu.hpp, all host threads) of the corpus described by `plans` (objects with
    burst_starts / burst_payloads / freq_offset_hz / seed, i.e. sameold_b200.synth.StreamPlan): int16 [len(plans),
    n_samples].  Same signal model as the device generator; used by bench.py's reference arm so that arm loads no GPU
    code."""
    n = len(plans)
    begin = np.zeros(n + 1, np.uint32)
    bursts, blobs, off = [], [], 0
    for i, p in enumerate(plans):
        for s, b in zip(p.burst_starts, p.burst_payloads):
            bursts.append((s, off, len(b)))
            blobs.append(b)
            off += len(b)
        begin[i + 1] = len(bursts)
    barr = (_CpuBurst * max(len(bursts), 1))()
    for k, (s, o, l) in enumerate(bursts):
        barr[k].start_sample, barr[k].byte_offset, barr[k].n_bytes = s, o, l
    data = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8).copy()
    foff = np.array([p.freq_offset_hz for p in plans], np.float32)
    seeds = np.array([p.seed for p in plans], np.uint32)
    out = np.empty((n, n_samples), np.int16)
    _lib().oracle_synth_generate(out.ctypes.data, n, n_samples, n_samples, rate, begin.ctypes.data, barr, data.ctypes.data,
                                 off, foff.ctypes.data, seeds.ctypes.data, amplitude, noise_sigma, int(n_threads))
    return out


def load_golden_recording(name):
    """int16 samples of one of the reference's sample/ recordings (tests/golden/<name>.22050.s16le.bin.gz)."""
    with gzip.open(os.path.join(GOLDEN_DIR, f"{name}.22050.s16le.bin.gz"), "rb") as f:
        return np.frombuffer(f.read(), dtype="<i2").copy()
