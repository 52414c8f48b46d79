"""Synthetic SAME corpus: host-side schedule + device-side waveform generation.

Tooling for bench.py and the GPU tests — not part of the receiver path.  Per stream (all values a deterministic
function of `seed_base + stream_id`, independent of how many streams are generated):

  * one SAME event = 3 header bursts + 3 EOM bursts; each burst = 16 x 0xAB preamble + payload, 1.0 s gaps,
    first burst at U(1,20) s, EOMs U(5,20) s after the last header burst;
  * header `ZCZC-ORG-EEE-PSSCCC(x1..31)+TTTT-JJJHHMM-LLLLLLLL-` drawn from small tables;
  * mark/space tones offset by U(-5,+5) Hz, amplitude 16384 (as receiver.rs:629), AWGN sigma = 11585/sqrt(10)
    (10 dB SNR relative to burst RMS) over the whole stream.

The waveform itself is produced on the GPU by same_synth_generate (csrc/same_synth.cu).
"""
import ctypes as C
from dataclasses import dataclass
from typing import List

import numpy as np

from . import _lib

SEED_BASE = 0x5A3E0000
PREAMBLE = bytes([0xAB] * 16)
AMPLITUDE = 16384.0
NOISE_SIGMA = 11585.0 / np.sqrt(10.0)
ORGS = ["EAS", "CIV", "WXR", "PEP"]
EVENTS = ["RWT", "RMT", "TOR", "SVR", "FFW", "EAN", "NPT", "CEM", "DMO", "WSW", "HUW", "SPS"]
CALLSIGN_CHARS = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789/"


class _SynthBurst(C.Structure):
    _fields_ = [("start_sample", C.c_double), ("byte_offset", C.c_uint32), ("n_bytes", C.c_uint32)]


@dataclass
class StreamPlan:
    header: str
    burst_starts: List[float]      # in samples
    burst_payloads: List[bytes]    # preamble + text
    freq_offset_hz: float
    seed: int


def _rng_for(stream_id: int, seed_base: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=seed_base + int(stream_id)))


def plan_stream(stream_id: int, rate: int = 22050, seconds: float = 60.0, seed_base: int = SEED_BASE) -> StreamPlan:
    r = _rng_for(stream_id, seed_base)
    nloc = int(r.integers(1, 32))
    locs = "-".join(f"{int(r.integers(0, 10))}{int(r.integers(0, 100)):02d}{int(r.integers(0, 1000)):03d}" for _ in range(nloc))
    call = "".join(CALLSIGN_CHARS[int(i)] for i in r.integers(0, len(CALLSIGN_CHARS), 8))
    hdr = (f"ZCZC-{ORGS[int(r.integers(0, len(ORGS)))]}-{EVENTS[int(r.integers(0, len(EVENTS)))]}-{locs}"
           f"+{int(r.integers(0, 24)):02d}{15 * int(r.integers(0, 4)):02d}-{int(r.integers(1, 366)):03d}"
           f"{int(r.integers(0, 24)):02d}{int(r.integers(0, 60)):02d}-{call}-")
    first = float(r.uniform(1.0, 20.0))
    gap = float(r.uniform(5.0, 20.0))
    foff = float(r.uniform(-5.0, 5.0))
    sym = rate / 520.83
    hb = PREAMBLE + hdr.encode("ascii")
    eb = PREAMBLE + b"NNNN"
    starts, payloads, t = [], [], first * rate
    for _ in range(3):
        starts.append(t); payloads.append(hb); t += len(hb) * 8 * sym + 1.0 * rate
    t += (gap - 1.0) * rate
    for _ in range(3):
        starts.append(t); payloads.append(eb); t += len(eb) * 8 * sym + 1.0 * rate
    total = seconds * rate
    keep = [(s, p) for s, p in zip(starts, payloads) if s + len(p) * 8 * sym < total]
    return StreamPlan(hdr, [k[0] for k in keep], [k[1] for k in keep], foff, (seed_base + int(stream_id)) & 0xFFFFFFFF)


def plan_long_stream(hours: float = 24.0, rate: int = 22050, seed: int = 0x24C0FFEE) -> StreamPlan:
    """ONE continuous stream, a SAME event (3 header + 3 EOM bursts, 1 s gaps, EOMs U(5,20) s after
    the header) every U(10,60) minutes, AWGN over the whole length."""
    r = np.random.Generator(np.random.Philox(key=seed))
    total = hours * 3600.0 * rate
    sym = rate / 520.83
    starts, payloads = [], []
    t = float(r.uniform(10.0, 60.0)) * 60.0 * rate
    k = 0
    first_header = ""
    while True:
        ev = plan_stream(100000 + k, rate, 120.0)      # header text / payloads of one event (its own start is ignored)
        first_header = first_header or ev.header
        hb, eb = PREAMBLE + ev.header.encode("ascii"), PREAMBLE + b"NNNN"
        gap = float(r.uniform(5.0, 20.0))
        t0, ss, pp = t, [], []
        for _ in range(3):
            ss.append(t0); pp.append(hb); t0 += len(hb) * 8 * sym + 1.0 * rate
        t0 += (gap - 1.0) * rate
        for _ in range(3):
            ss.append(t0); pp.append(eb); t0 += len(eb) * 8 * sym + 1.0 * rate
        if t0 + rate >= total:
            break
        starts += ss; payloads += pp
        t = t0 + float(r.uniform(10.0, 60.0)) * 60.0 * rate
        k += 1
    return StreamPlan(first_header, starts, payloads, float(r.uniform(-5.0, 5.0)), seed & 0xFFFFFFFF)


def plan_corpus(n_streams: int, rate: int = 22050, seconds: float = 60.0, first_stream: int = 0,
                seed_base: int = SEED_BASE) -> List[StreamPlan]:
    return [plan_stream(first_stream + i, rate, seconds, seed_base) for i in range(n_streams)]


class DeviceCorpus:
    """The burst tables of a list of plans, packed once; `generate` renders any time window of every stream."""

    def __init__(self, plans: List[StreamPlan], rate: int = 22050, device: int = 0, amplitude: float = AMPLITUDE,
                 noise_sigma: float = NOISE_SIGMA):
        self.lib = _lib.load()
        self.n, self.rate, self.device, self.amplitude, self.noise_sigma = len(plans), rate, device, amplitude, noise_sigma
        self.begin = np.zeros(self.n + 1, np.uint32)
        bursts, blobs, off = [], [], 0
        for i, p in enumerate(plans):
            for s, b in zip(p.burst_starts, p.burst_payloads):
                bursts.append((s, off, len(b)))
                blobs.append(b)
                off += len(b)
            self.begin[i + 1] = len(bursts)
        self.nbursts, self.nbytes = len(bursts), off
        self.barr = (_SynthBurst * max(len(bursts), 1))()
        for k, (s, o, l) in enumerate(bursts):
            self.barr[k].start_sample, self.barr[k].byte_offset, self.barr[k].n_bytes = s, o, l
        self.data = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8).copy()
        self.foff = np.array([p.freq_offset_hz for p in plans], np.float32)
        self.seeds = np.array([p.seed for p in plans], np.uint32)

    def generate(self, d_ptr: int, stride: int, n_samples: int, first_sample: int = 0):
        """d_ptr[stream * stride + k] = sample first_sample + k of each stream, k < n_samples (first_sample % 8 == 0)."""
        err = C.create_string_buffer(256)
        rc = self.lib.same_synth_generate(self.device, C.c_void_p(d_ptr), self.n, stride, int(first_sample), n_samples,
                                          self.rate, self.begin.ctypes.data, self.barr, self.nbursts,
                                          self.data.ctypes.data, self.nbytes, self.foff.ctypes.data,
                                          self.seeds.ctypes.data, self.amplitude, self.noise_sigma, err)
        if rc != 0:
            raise RuntimeError(f"same_synth_generate failed: {err.value.decode()}")


def generate_on_device(plans: List[StreamPlan], d_ptr: int, stride: int, n_samples: int, rate: int = 22050,
                       device: int = 0, amplitude: float = AMPLITUDE, noise_sigma: float = NOISE_SIGMA,
                       first_sample: int = 0):
    """Fill device memory d_ptr[stream * stride + k] (int16) with samples [first_sample, first_sample + n_samples) of
    the corpus described by `plans`."""
    DeviceCorpus(plans, rate, device, amplitude, noise_sigma).generate(d_ptr, stride, n_samples, first_sample)


def render_numpy(plan: StreamPlan, n_samples: int, rate: int = 22050, amplitude: float = AMPLITUDE,
                 noise_sigma: float = NOISE_SIGMA) -> np.ndarray:
    """CPU rendering of one stream (same signal model, numpy noise — NOT bit-identical to the device generator).
    Used by CPU-only tests; every decoder under test is always fed the identical int16 samples."""
    ts = rate / 520.83
    n = np.arange(n_samples, dtype=np.float64)
    sig = np.zeros(n_samples, np.float64)
    fm, fs = 2083.3 + plan.freq_offset_hz, 1562.5 + plan.freq_offset_hz
    for start, payload in zip(plan.burst_starts, plan.burst_payloads):
        bits = np.unpackbits(np.frombuffer(payload, np.uint8), bitorder="little").astype(np.int64)
        i0, i1 = int(np.ceil(start)), min(n_samples, int(np.ceil(start + len(bits) * ts)))
        if i1 <= i0:
            continue
        t = n[i0:i1] - start
        k = np.minimum((t / ts).astype(np.int64), len(bits) - 1)
        km = np.concatenate([[0], np.cumsum(bits)])[k]
        f = np.where(bits[k] == 1, fm, fs)
        cycles = (fm * km + fs * (k - km)) / 520.83 + f * (t - k * ts) / rate
        sig[i0:i1] = amplitude * np.cos(2.0 * np.pi * (cycles - np.floor(cycles)))
    noise = np.random.Generator(np.random.Philox(key=plan.seed)).standard_normal(n_samples) * noise_sigma
    return np.clip(np.rint(sig + noise), -32768, 32767).astype(np.int16)
