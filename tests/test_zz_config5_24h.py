"""BASELINE config 5 at full length: ONE continuous 24 h stream (1 905 120 000 samples, 3.81 GB) with sparse SAME events,
decoded by a one-stream engine in 10-minute chunks, against the CPU oracle event for event.

A single stream is strictly sequential in its timing loop (receiver.rs:243); everything in front of the loop runs
time-parallel in the engine's long-stream path (same_long.cu, DESIGN.md §4: speculative block AGC with bitwise-verified
hand-over, dense matched filters, bursts by the ordinary kernels).  The test pins parity over 24 h of state carried
across 144 submits and reports the single-stream realtime factor.  Runs last (file name): the longest test (~1 min).
SAME_TEST_HOURS overrides the length (e.g. 2 for a quick check)."""
import os
import time

import numpy as np
import pytest

import sameold_b200 as sb
from oracle import Oracle
from oracle.pyoracle import OracleConfig
from sameold_b200 import synth

pytestmark = pytest.mark.gpu


def test_single_24h_stream_vs_oracle():
    import torch
    assert torch.cuda.is_available()
    hours = float(os.environ.get("SAME_TEST_HOURS", "24"))
    rate = 22050
    n = int(hours * 3600 * rate)
    plan = synth.plan_long_stream(hours, rate)
    buf = torch.empty(((n + 7) // 8 * 8,), dtype=torch.int16, device="cuda")
    synth.DeviceCorpus([plan], rate).generate(buf.data_ptr(), buf.numel(), n)
    b = sb.SameReceiverBuilder.samedec(rate)
    rx = b.build_batch(1)
    step = 600 * rate
    got = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for lo in range(0, n, step):
        k = min(step, n - lo)
        rx.submit_device(buf.data_ptr() + 2 * lo, k, np.zeros(1, np.uint64), np.array([k], np.uint32))
        rx.sync()
        got.extend(rx.drain())
    gpu_s = time.perf_counter() - t0
    assert rx.input_sample_counters()[0] == n
    host = buf[:n].cpu().numpy()
    c, ocfg = b.config(), OracleConfig()
    for name, _ in OracleConfig._fields_:
        setattr(ocfg, name, getattr(c, name))
    o = Oracle(ocfg)
    t0 = time.perf_counter()
    o.process_s16(host)
    cpu_s = time.perf_counter() - t0
    want = o.events()
    n_msgs = sum(1 for e in want if e.kind in (18, 19))
    print(f"\nconfig 5: {hours:g} h stream, {len(plan.burst_starts)} bursts planned, {len(want)} events, {n_msgs} messages; "
          f"engine {gpu_s:.1f} s = {hours * 3600 / gpu_s:.0f}x realtime (long-stream path); oracle {cpu_s:.1f} s = "
          f"{hours * 3600 / cpu_s:.0f}x realtime (one core)")
    assert n_msgs >= 2 * (len(plan.burst_starts) // 6) - 2, "nearly every event yields a header and an EOM"
    assert [e.key() for e in got] == [e.key() for e in want]
