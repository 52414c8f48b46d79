"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on identical int16 inputs.

Bar (BASELINE.json north_star): link + transport events bit-exact — kind, input_sample_counter, symbol count, burst /
message bytes, parity and voting counts — and soft symbols bit-equal (the chain is chaotic: a tolerance would be
meaningless; the 1e-4 relative tolerance of the north star is asserted as well).
Nothing here reads /root/reference: recordings come from tests/golden/.
"""
import os

import numpy as np
import pytest

import sameold_b200 as sb
from oracle import GOLDEN_DIR, Oracle, decode_batch_events, load_golden_recording
from oracle.pyoracle import OracleConfig
from sameold_b200 import synth

pytestmark = pytest.mark.gpu

NAMES = ["long_message", "npt", "two_and_two"]


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def oracle_cfg_from(builder) -> OracleConfig:
    c, o = builder.config(), OracleConfig()
    for name, _ in OracleConfig._fields_:
        setattr(o, name, getattr(c, name))
    return o


def expected_lines(name):
    with open(os.path.join(GOLDEN_DIR, f"{name}.22050.s16le.txt")) as f:
        return [l.rstrip("\n") for l in f if not l.startswith("+OK")]


def assert_events_equal(gpu_events, oracle_events, ctx=""):
    g = [e.key() for e in gpu_events]
    o = [e.key() for e in oracle_events]
    if g != o:
        for i, (a, b) in enumerate(zip(g, o)):
            assert a == b, f"{ctx}: event {i} differs\n  gpu    {a}\n  oracle {b}"
        assert len(g) == len(o), f"{ctx}: {len(g)} gpu events vs {len(o)} oracle events"


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 2: the three sample/ recordings as one ragged 3-stream batch
# ---------------------------------------------------------------------------------------------------------------------
def test_golden_recordings_batch_bit_exact():
    _torch()
    recs = [load_golden_recording(n) for n in NAMES]
    b = sb.SameReceiverBuilder.samedec(22050)
    rx = b.build_batch(3)
    rx.enable_soft_trace(16384)
    evs = rx.process(recs)
    traces = [rx.read_soft_trace(s) for s in range(3)]
    tail = rx.flush_samedec()
    for s, name in enumerate(NAMES):
        o = Oracle(oracle_cfg_from(b))
        o.enable_trace()
        o.process_s16(recs[s])
        n_before_flush = len(o.events())
        ot = o.soft_trace()
        o.flush_samedec()
        oe = o.events()
        assert_events_equal(evs[s], oe[:n_before_flush], f"{name} (input)")
        assert_events_equal(evs[s] + tail[s], oe, f"{name} (input + EOF flush)")
        # messages == the reference's golden stdout
        msgs = [str(e.message_ok()) for e in evs[s] + tail[s] if e.message_ok() is not None]
        assert msgs == expected_lines(name)
        # soft symbols: bit-equal (and therefore within the stated 1e-4 relative tolerance)
        t = traces[s]
        assert len(t) == len(ot)
        assert np.array_equal(t["sample"], ot["sample"])
        assert np.array_equal(t["zero"].view(np.uint32), ot["zero"].view(np.uint32))
        assert np.array_equal(t["sym"].view(np.uint32), ot["sym"].view(np.uint32))
        assert np.allclose(t["sym"], ot["sym"], rtol=1e-4, atol=0)
    assert np.array_equal(rx.input_sample_counters()[:3] >= np.array([len(r) for r in recs]), [True] * 3)


def test_decode_samedec_matches_golden_text():
    _torch()
    rx = sb.SameReceiverBuilder.samedec(22050).build_batch(3)
    out = rx.decode_samedec([load_golden_recording(n) for n in NAMES])
    assert out == [expected_lines(n) for n in NAMES]


def test_chunked_submit_equals_whole():
    """State is resident between submits: arbitrary ragged chunking gives identical events (receiver.rs:233-274)."""
    _torch()
    recs = [load_golden_recording(n) for n in NAMES]
    b = sb.SameReceiverBuilder.samedec(22050)
    whole = b.build_batch(3).process(recs)
    rx = b.build_batch(3)
    rng = np.random.default_rng(7)
    pos = [0, 0, 0]
    got = [[], [], []]
    while any(p < len(r) for p, r in zip(pos, recs)):
        chunks = []
        for s in range(3):
            n = int(rng.integers(0, 60000)) if rng.random() > 0.1 else 0   # some streams skip a round
            chunks.append(recs[s][pos[s]:pos[s] + n])
            pos[s] += len(chunks[-1])
        for s, e in enumerate(rx.process(chunks)):
            got[s].extend(e)
    for s in range(3):
        assert_events_equal(got[s], whole[s], f"stream {s}")


def test_tiny_chunks_cross_every_boundary():
    """1..50-sample chunks: the sample clock, TED parity and byte clock all resume mid-period."""
    _torch()
    rec = load_golden_recording("npt")[:60000]
    b = sb.SameReceiverBuilder.samedec(22050)
    o = Oracle(oracle_cfg_from(b))
    o.process_s16(rec)
    rx = b.build_batch(1)
    rng = np.random.default_rng(3)
    got, i = [], 0
    # first 3000 samples in tiny chunks, then the burst region in moderate chunks
    while i < len(rec):
        n = int(rng.integers(1, 50)) if i < 3000 else int(rng.integers(200, 5000))
        got.extend(rx.process([rec[i:i + n]])[0])
        i += n
    assert_events_equal(got, o.events(), "tiny chunks")


def test_library_default_config_and_event_order():
    """Library defaults (AGC limits [0, 1e6], builder.rs:55) on a clean test burst: the event ORDER of the reference's
    test_iter_events (receiver.rs:651-671) and bit-exact parity with the oracle."""
    _torch()
    plan = synth.StreamPlan("", [22050.0], [synth.PREAMBLE + b"ZCZC-EAS-DMO-372088-091724+0000-0001122-NOCALL00-"], 0.0, 1)
    x = synth.render_numpy(plan, 5 * 22050, noise_sigma=0.0)
    b = sb.SameReceiverBuilder(22050).with_timing_max_deviation(0.01)
    evs = b.build_batch(1).process([x])[0]
    o = Oracle(oracle_cfg_from(b))
    o.process_s16(x)
    assert_events_equal(evs, o.events(), "library defaults")
    kinds = [e.kind for e in evs]
    assert kinds == [1, 2, 3, 17, 0], kinds  # Searching, Reading, Burst, Assembling, NoCarrier
    assert evs[2].data.startswith(b"ZCZC-EAS-DMO-372088-091724+0000-0001122-NOCALL00-")


def test_derived_constants_bit_equal():
    _torch()
    for rate in (22050, 44100, 48000, 11025):
        b = sb.SameReceiverBuilder.samedec(rate)
        d = b.build_batch(1).derived()
        o = Oracle.derived(oracle_cfg_from(b))
        for k, v in o.items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(d[k].view(np.uint32), v.view(np.uint32)), (rate, k)
            elif isinstance(v, float):
                assert np.float32(d[k]).view(np.uint32) == np.float32(v).view(np.uint32), (rate, k)
            else:
                assert d[k] == v, (rate, k)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 3 (scaled to what the oracle finishes in seconds): device-generated synthetic corpus
# ---------------------------------------------------------------------------------------------------------------------
def _device_corpus(n_streams, seconds, rate=22050, first=0):
    torch = _torch()
    n = int(seconds * rate)
    stride = (n + 7) // 8 * 8
    buf = torch.empty((n_streams, stride), dtype=torch.int16, device="cuda")
    plans = synth.plan_corpus(n_streams, rate, seconds, first_stream=first)
    synth.generate_on_device(plans, buf.data_ptr(), stride, n, rate)
    return buf, plans, n, stride


def test_synthetic_corpus_bit_exact_vs_oracle():
    _torch()
    ns, secs = 96, 45.0      # 3 warps, ragged tail warp excluded on purpose below
    buf, plans, n, stride = _device_corpus(ns, secs)
    host = buf.cpu().numpy()
    b = sb.SameReceiverBuilder.samedec(22050)
    rx = b.build_batch(ns)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
    rx.submit_device(buf.data_ptr(), ns * stride, offsets, np.full(ns, n, np.uint32))
    rx.sync()
    evs = rx.drain_by_stream()
    decoded = 0
    for s in range(ns):
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(host[s, :n])
        assert_events_equal(evs[s], o.events(), f"synthetic stream {s}")
        msgs = [str(e.message_ok()) for e in evs[s] if e.message_ok() is not None]
        if plans[s].header in msgs:
            decoded += 1
    # the corpus is decodable: at 10 dB SNR nearly every header comes through
    assert decoded >= int(0.9 * ns), decoded


def test_device_and_host_submit_agree_and_ragged_lengths():
    _torch()
    ns = 40                  # not a multiple of 32: exercises the padded lanes
    buf, plans, n, stride = _device_corpus(ns, 12.0, first=1000)
    host = buf.cpu().numpy()
    lengths = np.array([n - 997 * (s % 7) for s in range(ns)], np.uint32)
    lengths[5] = 0           # a stream that receives nothing
    lengths[6] = 1
    b = sb.SameReceiverBuilder.samedec(22050)
    rx_d = b.build_batch(ns)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
    rx_d.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
    rx_d.sync()
    ev_d = rx_d.drain_by_stream()
    rx_h = b.build_batch(ns)
    ev_h = rx_h.process([host[s, :lengths[s]] for s in range(ns)])
    assert np.array_equal(rx_d.input_sample_counters(), lengths.astype(np.uint64))
    for s in range(ns):
        assert_events_equal(ev_d[s], ev_h[s], f"stream {s} device vs host submit")
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(host[s, :lengths[s]])
        assert_events_equal(ev_h[s], o.events(), f"stream {s}")


def test_fast_and_generic_kernels_agree():
    """The 22050 Hz fast kernels (integer DC blocker, packed FFMA2 matched filter, 16-byte loads; warp-specialised
    producer/consumer, five-warp pipelined and single-warp forms) and the rate-generic kernel (literal f32 recursion) must give identical
    events AND identical resident state: chunks alternate between the three kernels mid-stream, with ragged, odd and
    unaligned chunk lengths."""
    _torch()
    recs = [load_golden_recording(n) for n in NAMES] + [synth.render_numpy(synth.plan_stream(5, seconds=20.0), 20 * 22050)]
    b = sb.SameReceiverBuilder.samedec(22050)
    ref = b.build_batch(len(recs))
    ref.set_option("kernel", 1)
    want = ref.process(recs)
    rx = b.build_batch(len(recs))
    rng = np.random.default_rng(11)
    pos = [0] * len(recs)
    got = [[] for _ in recs]
    k = 0
    while any(p < len(r) for p, r in zip(pos, recs)):
        rx.set_option("kernel", 1 + k % 6)   # 1 generic, 2 single-warp fast, 3 pipelined, 4 three-warp, 5 split (front end + tile-fed), 6 look-ahead
        k += 1
        chunks = []
        for s_ in range(len(recs)):
            n = int(rng.integers(1, 40000))
            chunks.append(recs[s_][pos[s_]:pos[s_] + n])
            pos[s_] += len(chunks[-1])
        for s_, e in enumerate(rx.process(chunks)):
            got[s_].extend(e)
    for s_ in range(len(recs)):
        assert_events_equal(got[s_], want[s_], f"stream {s_} alternating kernels")
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(recs[s_])
        assert_events_equal(want[s_], o.events(), f"stream {s_} generic kernel vs oracle")


@pytest.mark.parametrize("kernel", [2, 3, 4, 5, 6])
def test_each_fast_kernel_matches_oracle(kernel):
    """Every fast-kernel flavour on its own, whole streams in one submit and in 3 s chunks: golden recordings and
    synthetic streams with bursts against the oracle, event for event."""
    _torch()
    recs = [load_golden_recording(n) for n in NAMES]
    recs += [synth.render_numpy(synth.plan_stream(100 + i, seconds=45.0), 45 * 22050) for i in range(6)]
    b = sb.SameReceiverBuilder.samedec(22050)
    want = []
    for r in recs:
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(r)
        want.append(o.events())
    rx = b.build_batch(len(recs))
    rx.set_option("kernel", kernel)
    got = rx.process(recs)
    for s_ in range(len(recs)):
        assert_events_equal(got[s_], want[s_], f"kernel {kernel} stream {s_} one submit")
    rx = b.build_batch(len(recs))
    rx.set_option("kernel", kernel)
    got = [[] for _ in recs]
    step = 3 * 22050 + 7
    for lo in range(0, max(len(r) for r in recs), step):
        for s_, e in enumerate(rx.process([r[lo:lo + step] for r in recs])):
            got[s_].extend(e)
    for s_ in range(len(recs)):
        assert_events_equal(got[s_], want[s_], f"kernel {kernel} stream {s_} chunked")


def _canonical_raw(evs, pay):
    """Raw drained events in an order- and arena-independent form: records with the payload offset zeroed + the payload
    bytes concatenated in event order."""
    e = evs.copy()
    off, ln = e["data_offset"].astype(np.int64), e["data_len"].astype(np.int64)
    ln = np.where(e["kind"] == 3, np.minimum(ln, 1024), ln)     # burst payloads are capped at SAME_BURST_CAP
    e["data_offset"] = 0
    idx = np.repeat(off - np.concatenate(([0], np.cumsum(ln)[:-1])), ln) + np.arange(int(ln.sum()))
    return e, pay[idx]


@pytest.mark.parametrize("ns", [5000, 20001])
def test_large_batches_engine_policy_matches_generic_kernel(ns):
    """Batches beyond one block per SM (engine picks the three-warp kernel) and beyond four (single-warp kernel), with a
    ragged tail warp: same events as the rate-generic kernel on every stream, and as the oracle on a sample."""
    _torch()
    secs = 22.0
    buf, plans, n, stride = _device_corpus(ns, secs, first=7000)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
    lengths = (n - (np.arange(ns) % 97) * 5).astype(np.uint32)
    b = sb.SameReceiverBuilder.samedec(22050)
    got = []
    for kernel in (0, 1):
        rx = b.build_batch(ns)
        rx.set_option("kernel", kernel)
        rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
        rx.sync()
        got.append(_canonical_raw(*rx.drain_raw()))
        del rx
    assert got[0][0].size > ns          # the corpus has bursts
    assert np.array_equal(got[0][0], got[1][0])
    assert np.array_equal(got[0][1], got[1][1])
    # oracle on a sample of streams (first, last, a few in between)
    rx = b.build_batch(ns)
    rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
    rx.sync()
    by_stream = rx.drain_by_stream()
    for s_ in (0, 31, 32, ns // 2, ns - 33, ns - 1):
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(buf[s_, :int(lengths[s_])].cpu().numpy())
        assert_events_equal(by_stream[s_], o.events(), f"stream {s_} of {ns}")


def _assert_raw_equal(got, want, ctx):
    """Whole event arrays (engine same_event records vs the oracle's, both canonicalised) must be identical; on a
    mismatch report the first differing stream/event."""
    (ge, gp), (we, wp) = got, want
    if ge.size == we.size and np.array_equal(ge, we) and np.array_equal(gp, wp):
        return
    n = min(ge.size, we.size)
    bad = np.nonzero(ge[:n] != we[:n])[0]
    i = int(bad[0]) if bad.size else n
    raise AssertionError(f"{ctx}: {ge.size} engine events vs {we.size} oracle events; first difference at event {i}:\n"
                         f"  engine {ge[i] if i < ge.size else None}\n  oracle {we[i] if i < we.size else None}")


def _rebase_seq(evs):
    """Per-stream sequence numbers counted from 0 (the engine's count from create/reset; the oracle batch export's
    always from 0) -- `evs` is sorted by (stream, seq)."""
    e = evs.copy()
    if e.size:
        first = np.concatenate(([True], e["stream"][1:] != e["stream"][:-1]))
        base = np.maximum.accumulate(np.where(first, np.arange(e.size), 0))
        e["seq"] = (np.arange(e.size) - base).astype(np.uint32)
    return e


def test_config3_full_size_every_stream_vs_oracle():
    """BASELINE config 3 at full size (4096 streams x 60 s, 10.8 GB of samples), BASELINE.md §3 row 3: the engine's
    events on EVERY stream are bit-identical to the CPU oracle's (kind, sample counter, symbol count, bytes, parity /
    voting counts); the engine's kernel for this batch size (pipelined) run three times, and every other kernel once
    (generic, single-warp, look-ahead single-warp, three-warp, split front end + tile-fed), produce the same
    event stream; every stream decodes its planned header."""
    _torch()
    ns, secs = 4096, 60.0
    buf, plans, n, stride = _device_corpus(ns, secs)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
    lengths = np.full(ns, n, np.uint32)
    b = sb.SameReceiverBuilder.samedec(22050)
    # the oracle on all 4096 x 60 s, one receiver per stream, every host core
    host = buf.cpu().numpy()
    oe, op, osecs = decode_batch_events(oracle_cfg_from(b), host[:, :n], os.cpu_count() or 1)
    want = _canonical_raw(oe, op)
    del host
    ref = None
    for kernel in (0, 0, 0, 1, 2, 4, 5, 6):
        rx = b.build_batch(ns)
        rx.set_option("kernel", kernel)
        rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
        rx.sync()
        evs, pay = _canonical_raw(*rx.drain_raw())
        del rx
        if ref is None:
            ref = (evs, pay)
            _assert_raw_equal((evs, pay), want, "config 3, all 4096 streams, engine vs oracle")
            som = evs[evs["kind"] == 18]
            assert np.array_equal(np.unique(som["stream"]), np.arange(ns)), "every stream decodes its header"
        else:
            assert np.array_equal(evs, ref[0]), f"kernel {kernel}: events differ"
            assert np.array_equal(pay, ref[1]), f"kernel {kernel}: payload differs"
    # the decoded header text is the planned one (spot check, payload of the first SOM event of a few streams)
    rx = b.build_batch(ns)
    rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
    rx.sync()
    by_stream = rx.drain_by_stream()
    for s_ in (0, 1, 777, 4095):
        msgs = [e for e in by_stream[s_] if e.kind == 18]
        assert msgs and bytes(msgs[0].data).decode("ascii") == plans[s_].header, f"stream {s_}"


def test_config4_65536_streams_time_chunked_two_shards_vs_oracle():
    """BASELINE config 4: 65 536 synthetic 60 s streams (173 GB of samples: more than one GPU holds), streamed through
    in 10 s time-chunks with the receiver state resident, as two contiguous 32 768-stream shards (one engine + one
    host thread each, same_multi_*).  Every 64th stream (1024 streams, all 60 s) is compared with the oracle event
    for event; chunked == whole follows from receiver.rs:233-274."""
    torch = _torch()
    ns, secs, chunk_s, every = 65536, 60.0, 10.0, 64
    rate = 22050
    n = int(secs * rate)
    cn = (int(chunk_s * rate) + 7) // 8 * 8          # chunk starts must be multiples of 8 samples (generator window)
    nchunks = (n + cn - 1) // cn
    plans = synth.plan_corpus(ns, rate, secs)
    corpus = synth.DeviceCorpus(plans, rate)
    buf = torch.empty((ns, cn), dtype=torch.int16, device="cuda")
    b = sb.SameReceiverBuilder.samedec(rate)
    rx = b.build_multi(ns, [0, 0])
    assert [(f, c) for _, f, c in rx.shards()] == [(0, 32768), (32768, 32768)]
    host = torch.empty((ns // every, n), dtype=torch.int16, pin_memory=False)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(cn)
    import ctypes as C
    lib = rx._lib
    evs_all, pay_all, pay_base = [], [], 0
    for c in range(nchunks):
        w = min(cn, n - c * cn)                      # the last chunk is shorter
        lengths = np.full(ns, w, np.uint32)
        corpus.generate(buf.data_ptr(), cn, w, first_sample=c * cn)
        host[:, c * cn:c * cn + w] = buf[::every, :w].cpu()
        for i, (_dev, first, count) in enumerate(rx.shards()):   # device-resident chunk: each shard's engine takes its rows
            eng = C.c_void_p(lib.same_multi_engine(rx._h, i))
            rc = lib.same_engine_submit_s16_device(eng, C.c_void_p(buf.data_ptr() + first * cn * 2), count * cn,
                                                   offsets[:count].ctypes.data, lengths[:count].ctypes.data)
            assert rc == 0
        rx.sync()
        e, p = rx.drain_raw()
        e = e.copy(); e["data_offset"] += pay_base
        evs_all.append(e); pay_all.append(p.copy()); pay_base += p.size
    assert np.array_equal(rx.input_sample_counters(), np.full(ns, n, np.uint64))
    evs = np.concatenate(evs_all); pay = np.concatenate(pay_all)
    order = np.lexsort((evs["seq"], evs["stream"]))          # per-chunk batches -> global (stream, occurrence) order
    evs = evs[order]
    assert int((evs["kind"] == 18).sum()) >= int(0.95 * ns), "the corpus is decodable"
    sel = evs[evs["stream"] % every == 0]
    sel["stream"] //= every
    got = _canonical_raw(_rebase_seq(sel), pay)
    oe, op, _ = decode_batch_events(oracle_cfg_from(b), host.numpy(), os.cpu_count() or 1)
    _assert_raw_equal(got, _canonical_raw(oe, op), "config 4, 1024 sampled streams, engine vs oracle")


def test_device_and_host_event_sort_agree():
    """Events come back per stream in order of occurrence; big batches are ordered on the device, small or accumulated
    ones on the host.  Both must give the same array, element for element, also across two submits before a drain."""
    _torch()
    ns = 3000
    buf, plans, n, stride = _device_corpus(ns, 24.0, first=9000)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
    lengths = np.full(ns, n, np.uint32)
    half = np.full(ns, n // 2, np.uint32)
    b = sb.SameReceiverBuilder.samedec(22050)
    out = []
    for device_sort, split in ((1, False), (0, False), (1, True)):
        rx = b.build_batch(ns)
        rx.set_option("device_sort", device_sort)
        if split:   # two collects before the drain: the second batch is merged on the host
            rx.submit_device(buf.data_ptr(), ns * stride, offsets, half)
            rx.sync()
            rx.submit_device(buf.data_ptr(), ns * stride, offsets + np.uint64(n // 2), lengths - half)
            rx.sync()
        else:
            rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
            rx.sync()
        evs, pay = rx.drain_raw()
        assert evs.size > 4096
        key = evs["stream"].astype(np.uint64) << np.uint64(32) | evs["seq"].astype(np.uint64)
        assert np.all(np.diff(key.astype(np.int64)) > 0), "sorted by (stream, occurrence), no duplicates"
        out.append(_canonical_raw(evs, pay))
    for other in out[1:]:
        assert np.array_equal(out[0][0], other[0]) and np.array_equal(out[0][1], other[1])


def test_long_single_stream_in_time_chunks():
    """Config 5 shape (one continuous stream, sparse bursts), shortened to 16 minutes: clusters of SAME bursts separated
    by minutes of noise, fed in 60 s chunks to a one-stream engine; events and messages must equal the oracle's."""
    _torch()
    rng = np.random.default_rng(5)
    parts = []
    for i in range(4):
        parts.append(synth.render_numpy(synth.plan_stream(300 + i, seconds=60.0), 60 * 22050))
        parts.append(np.clip(np.rint(rng.normal(0.0, 3663.0, 180 * 22050)), -32768, 32767).astype(np.int16))
    stream = np.concatenate(parts)
    b = sb.SameReceiverBuilder.samedec(22050)
    o = Oracle(oracle_cfg_from(b))
    o.process_s16(stream)
    want = o.events()
    assert sum(1 for e in want if e.is_message) >= 4
    rx = b.build_batch(1)
    got = []
    step = 60 * 22050
    for lo in range(0, len(stream), step):
        got.extend(rx.process([stream[lo:lo + step]])[0])
    assert_events_equal(got, want, "16-minute stream in 60 s chunks")
    assert rx.input_sample_counters()[0] == len(stream)


def test_long_stream_path_is_bit_exact_and_used():
    """One stream, chunks of >= 65536 samples take the long-stream path (same_long.cu): DC blocker, AGC and matched
    filters time-parallel (the AGC speculatively per block, every hand-over gain verified bitwise), the timing loop and
    everything after it sequential on one lane; bursts (AGC locked) by the ordinary tile-fed kernel.  Events must equal
    the oracle's and the ordinary kernels', across chunk boundaries that fall inside bursts, and the path must really
    have run (speculative passes and fallback spans counted)."""
    _torch()
    b = sb.SameReceiverBuilder.samedec(22050)
    rng = np.random.default_rng(17)
    streams = [load_golden_recording(n) for n in NAMES]
    parts = []
    for i in range(3):       # 3 SAME events with minutes of noise between them
        parts.append(synth.render_numpy(synth.plan_stream(700 + i, seconds=60.0), 60 * 22050))
        parts.append(np.clip(np.rint(rng.normal(0.0, 3663.0, 100 * 22050 + 977 * i)), -32768, 32767).astype(np.int16))
    streams.append(np.concatenate(parts))
    silence = np.zeros(20 * 22050, np.int16)                       # exact zeros: the gain runs into its clamp
    streams.append(np.concatenate([silence, streams[1], silence]))
    for si, x in enumerate(streams):
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(x)
        want = o.events()
        for chunking in ("whole", "ragged"):
            rx = b.build_batch(1)
            assert rx.get_option("long_stream") == 1
            got = []
            if chunking == "whole":
                got = rx.process([x])[0]
            else:
                i = 0
                while i < len(x):
                    k = int(rng.integers(65536, 400000))
                    got.extend(rx.process([x[i:i + k]])[0])
                    i += k
            assert_events_equal(got, want, f"long-stream path, stream {si}, {chunking}")
            assert rx.get_option("long_stream_passes") >= 1
            if any(e.kind == 3 for e in want):
                assert rx.get_option("long_stream_fallback_spans") >= 1
            assert rx.input_sample_counters()[0] == len(x)
        # and switching the path off gives the same events (ordinary pipelined kernel)
        rx = b.build_batch(1)
        rx.set_option("long_stream", 0)
        assert_events_equal(rx.process([x])[0], want, f"ordinary path, stream {si}")
        assert rx.get_option("long_stream_passes") == 0
    # a long chunk, then short ones (ordinary kernels), then a long one again: one resident state for all
    x = streams[3]
    o = Oracle(oracle_cfg_from(b))
    o.process_s16(x)
    rx = b.build_batch(1)
    got, i = [], 0
    for k in (300000, 1000, 31, 70000, 5000, 10 ** 9):
        got.extend(rx.process([x[i:i + k]])[0])
        i += k
    assert_events_equal(got, o.events(), "long and short chunks mixed")


def test_lane_sparse_warps_give_identical_results():
    """Small batches run with fewer streams per warp (latency-bound regime); the mapping must not change results."""
    _torch()
    ns = 70
    buf, plans, n, stride = _device_corpus(ns, 8.0, first=500)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
    lengths = np.array([n - 13 * s_ for s_ in range(ns)], np.uint32)
    b = sb.SameReceiverBuilder.samedec(22050)
    want = None
    for lanes in (32, 8, 4, 1):
        rx = b.build_batch(ns)
        rx.set_option("lanes_per_warp", lanes)
        rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
        rx.sync()
        got = rx.drain_by_stream()
        if want is None:
            want = got
            assert sum(len(g) for g in got) > 50
        else:
            for s_ in range(ns):
                assert_events_equal(got[s_], want[s_], f"lanes={lanes} stream {s_}")


def test_other_sample_rates_generic_kernel():
    """44.1 kHz (84 taps, DC length 32) and 48 kHz (92 taps, DC length 35: the DC blocker is no longer exact in f32 and
    must be evaluated sequentially in the reference's order) — SURVEY.md §8f N4."""
    _torch()
    for rate in (44100, 48000, 11025):
        plan = synth.plan_stream(77, rate, 30.0)
        x = synth.render_numpy(plan, int(30 * rate), rate)
        b = sb.SameReceiverBuilder.samedec(rate)
        evs = b.build_batch(1).process([x])[0]
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(x)
        assert_events_equal(evs, o.events(), f"rate {rate}")
        assert any(e.kind == 3 for e in evs), f"no burst decoded at {rate} Hz"


def test_non_default_configs():
    _torch()
    x = load_golden_recording("two_and_two")
    variants = [
        sb.SameReceiverBuilder.samedec(22050).without_adaptive_equalizer(),
        sb.SameReceiverBuilder.samedec(22050).with_adaptive_equalizer(sb.EqualizerBuilder().with_filter_order(8, 3).with_relaxation(0.1)),
        sb.SameReceiverBuilder.samedec(22050).with_timing_bandwidth(0.2, 0.02).with_preamble_max_errors(4).with_squelch_power(0.2, 0.1),
        sb.SameReceiverBuilder(22050).with_dc_blocker_length(0.03).with_frame_max_invalid(0).with_frame_prefix_max_errors(0),
    ]
    for i, b in enumerate(variants):
        evs = b.build_batch(1).process([x])[0]
        o = Oracle(oracle_cfg_from(b))
        o.process_s16(x)
        assert_events_equal(evs, o.events(), f"variant {i}")


# ---------------------------------------------------------------------------------------------------------------------
# Boundary behaviour
# ---------------------------------------------------------------------------------------------------------------------
def test_empty_and_reset_and_errors():
    _torch()
    b = sb.SameReceiverBuilder.samedec(22050)
    rx = b.build_batch(2)
    assert rx.process([np.zeros(0, np.int16), None]) == [[], []]
    assert list(rx.input_sample_counters()) == [0, 0]
    rec = load_golden_recording("npt")
    first = rx.process([rec, rec[:50000]])
    # reset stream 0 only: it starts over with AGC gain 1.0 (agc.rs:61), stream 1 keeps going
    rx.reset([0])
    assert list(rx.input_sample_counters()) == [0, 50000]
    second = rx.process([rec, rec[50000:]])
    o = Oracle(oracle_cfg_from(b))
    o.process_s16(rec)
    assert_events_equal(first[0], o.events(), "stream 0 first pass")
    assert_events_equal(first[1] + second[1], o.events(), "stream 1 across the reset of stream 0")
    o.reset()
    o.process_s16(rec)
    assert_events_equal(second[0], o.events(), "stream 0 after reset()")
    # invalid configurations are rejected like the reference's panics
    with pytest.raises(sb.SameEngineError) as ei:
        sb.SameReceiverBuilder(22050).with_dc_blocker_length(0.0).build_batch(1)
    assert ei.value.code == 2
    with pytest.raises(ValueError):
        rx.submit([rec])


def test_single_stream_facade_and_flush():
    """SameReceiver.iter_messages / flush (receiver.rs:155-161, 216-224), incl. the forced EOM after 135 s."""
    _torch()
    rec = load_golden_recording("long_message")
    rx = sb.SameReceiverBuilder.samedec(22050).build()
    assert rx.input_rate() == 22050
    assert list(rx.iter_messages(rec)) == []          # the header is still pending at EOF
    m = rx.flush()
    assert m is not None and m.is_start and str(m) == expected_lines("long_message")[0]
    assert m.voting_byte_count == 252 and m.parity_error_count == 0
    # flush() stopped AT the message sample, exactly like the reference
    o = Oracle.samedec()
    o.process_s16(rec)
    o.flush_samedec()
    som = [e for e in o.events() if e.kind == 18][0]
    assert rx.input_sample_counter() == som.sample
    assert rx.flush() is None                          # next 4 s of zeros: nothing
    # 135 s of silence later the receiver forces an EndOfMessage (receiver.rs:300-309)
    msgs = list(rx.iter_messages(np.zeros(136 * 22050, np.int16)))
    assert [str(x) for x in msgs] == ["NNNN"]
    rx.reset()
    assert rx.input_sample_counter() == 0


def test_snapshot_restore_is_clone():
    _torch()
    rec = load_golden_recording("two_and_two")
    rx = sb.SameReceiverBuilder.samedec(22050).build_batch(1)
    a = rx.process([rec[:80000]])[0]
    snap = rx.snapshot()
    b1 = rx.process([rec[80000:]])[0]
    rx.restore(snap)
    b2 = rx.process([rec[80000:]])[0]
    rx.free_snapshot(snap)
    assert_events_equal(b1, b2, "after restore")
    assert len(a) + len(b1) > 10


def test_event_overflow_is_reported_and_stays_consistent():
    """Arena too small: sync reports SAME_ERR_EVENT_OVERFLOW once; what is then drained is self-consistent
    (no payload offset past the arena; lost payloads flagged, lost events counted); the engine keeps working."""
    _torch()
    rec = load_golden_recording("npt")
    b = sb.SameReceiverBuilder.samedec(22050)
    want = b.build_batch(1).process([rec])[0]
    rx = b.build_batch(1)
    rx.set_event_capacity(4, 48)          # room for 4 events and one 45-byte burst
    with pytest.raises(sb.SameEngineError) as ei:
        rx.process([rec])
    assert ei.value.code == 5
    evs, pay = rx.drain_raw()
    assert evs.size == 4 and pay.size <= 48
    stored = np.where(evs["kind"] == 3, np.minimum(evs["data_len"], 1024), evs["data_len"])
    assert np.all(evs["data_offset"].astype(np.int64) + stored <= pay.size)
    for g, w in zip(sb.SameBatchReceiver.events_from_raw(evs, pay), want):
        assert (g.kind, g.sample, g.symbol_count) == (w.kind, w.sample, w.symbol_count)
        assert g.data == w.data or (g.flags & 2 and g.data == b"")
    lost_ev, lost_pay = rx.lost_events()
    assert lost_ev == len(want) - 4
    # with a big enough arena the same engine carries on: chunked == whole still holds for what follows
    rx.set_event_capacity(65536, 1 << 20)
    o = Oracle(oracle_cfg_from(b))
    o.process_s16(rec)
    o.process_s16(rec)
    n_first = len(want)
    assert_events_equal(rx.process([rec])[0], o.events()[n_first:], "after the overflow")
    # payload arena smaller than one burst: the event arrives without bytes and says so
    rx2 = b.build_batch(1)
    rx2.set_event_capacity(1024, 16)
    with pytest.raises(sb.SameEngineError):
        rx2.process([rec])
    evs, pay = rx2.drain_raw()
    bursts = evs[evs["kind"] == 3]
    assert bursts.size == 3 and np.all(bursts["data_len"] == 0) and np.all(bursts["flags"] & 2)
    assert rx2.lost_events()[1] >= 3


def test_f32_ingest_matches_oracle():
    """The reference's own item type (iter_events<Item = f32>, receiver.rs:119-130; lib.rs:78-79 documents f32 PCM):
    normalised, non-integer samples with the library-default AGC limits [0, 1e6] (builder.rs:55) and with samedec's
    limits on unnormalised floats, whole and chunked, bit-exact against the oracle's f32 entry."""
    _torch()
    rng = np.random.default_rng(21)
    recs = [load_golden_recording(n) for n in NAMES]
    cases = [(sb.SameReceiverBuilder(22050), np.float32(1.0 / 32768.0)),
             (sb.SameReceiverBuilder.samedec(22050), np.float32(0.7301)),
             (sb.SameReceiverBuilder.samedec(44100), np.float32(0.25))]
    for b, scale in cases:
        if b.input_rate() == 22050:
            xs = [r.astype(np.float32) * scale for r in recs]
        else:
            plan = synth.plan_stream(78, b.input_rate(), 20.0)
            xs = [synth.render_numpy(plan, int(20 * b.input_rate()), b.input_rate()).astype(np.float32) * scale]
        want = []
        for x in xs:
            o = Oracle(oracle_cfg_from(b))
            o.process_f32(x)
            want.append(o.events())
        assert any(e.kind in (18, 19) for w in want for e in w)
        got = b.build_batch(len(xs)).process_f32(xs)
        for s_ in range(len(xs)):
            assert_events_equal(got[s_], want[s_], f"f32 x{scale} stream {s_} one submit")
        rx = b.build_batch(len(xs))
        got = [[] for _ in xs]
        pos = [0] * len(xs)
        while any(p < len(x) for p, x in zip(pos, xs)):
            chunks = []
            for s_ in range(len(xs)):
                k = int(rng.integers(1, 50000))
                chunks.append(xs[s_][pos[s_]:pos[s_] + k]); pos[s_] += len(chunks[-1])
            for s_, e in enumerate(rx.process_f32(chunks)):
                got[s_].extend(e)
        for s_ in range(len(xs)):
            assert_events_equal(got[s_], want[s_], f"f32 x{scale} stream {s_} chunked")


def test_f32_then_s16_on_one_engine_and_facade():
    """After an f32 chunk the DC-blocker windows may hold non-integers: the engine must stay on the literal f32
    recursion for later s16 chunks (until reset).  Also the single-stream facade with float arrays."""
    _torch()
    rec = load_golden_recording("two_and_two")
    b = sb.SameReceiverBuilder.samedec(22050)
    half = 70001
    xf = rec[:half].astype(np.float32) + np.float32(0.37)      # non-integer DC offset
    o = Oracle(oracle_cfg_from(b))
    o.process_f32(xf)
    o.process_s16(rec[half:])
    rx = b.build_batch(1)
    assert rx.get_option("kernel_selected") == 3
    got = rx.process_f32([xf])[0]
    assert rx.get_option("kernel_selected") == 1
    got += rx.process([rec[half:]])[0]
    assert_events_equal(got, o.events(), "f32 chunk then s16 chunk")
    rx.reset()
    assert rx.get_option("kernel_selected") == 3
    # facade: float arrays take the f32 entry, int16 arrays the s16 one
    one = sb.SameReceiverBuilder(22050).build()
    x = load_golden_recording("npt").astype(np.float32) / np.float32(32768.0)
    msgs = [str(m) for m in one.iter_messages(x)]
    assert msgs == expected_lines("npt")


def test_submit_2d_odd_columns_from_pinned_host_memory():
    """same_engine_submit_s16_2d with odd chunk widths and column starts (device rows are re-pitched to a
    multiple of 8 samples so the vector-load path stays on): chunked == whole, for the three fast kernels."""
    _torch()
    import ctypes as C
    from sameold_b200 import _lib
    lib = _lib.load()
    ns, n = 37, 21 * 22050 + 5
    stride = n + 3
    plans = synth.plan_corpus(ns, 22050, 21.0, first_stream=4242)
    hptr = lib.same_host_alloc(ns * stride * 2)
    assert hptr
    try:
        host = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_int16)), shape=(ns, stride))
        for s_, pl in enumerate(plans):
            host[s_, :n] = synth.render_numpy(pl, n)
        b = sb.SameReceiverBuilder.samedec(22050)
        want = b.build_batch(ns).process([host[s_, :n].copy() for s_ in range(ns)])
        assert sum(len(w) for w in want) > 150
        for kernel in (2, 3, 4, 6):
            rx = b.build_batch(ns)
            rx.set_option("kernel", kernel)
            got = [[] for _ in range(ns)]
            col = 0
            for width in (1, 7, 55125, 33333, 22051, n):
                w = min(width, n - col)
                if w <= 0:
                    break
                rx.submit_2d(hptr, stride, col, w)
                col += w
                rx.sync()
                for s_, e in enumerate(rx.drain_by_stream()):
                    got[s_].extend(e)
            assert col == n
            for s_ in range(ns):
                assert_events_equal(got[s_], want[s_], f"kernel {kernel} stream {s_} 2-D odd chunks")
    finally:
        lib.same_host_free(hptr)


def test_truncated_burst_longer_than_the_engine_buffer():
    """framing.rs:152-162 has no burst length cap; the engine keeps SAME_BURST_CAP = 1024 bytes and says so
    (SAME_EV_FLAG_TRUNCATED, data_len = true length).  Everything else -- sample indices, the transport layer, which
    only reads 268 bytes (assembler.rs:169) -- must still match the oracle."""
    _torch()
    text = b"ZCZC-" + (b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-+/ " * 40)[:1500]
    plan = synth.StreamPlan("", [11025.0], [synth.PREAMBLE + text], 1.5, 99)
    n = int(30 * 22050)
    x = synth.render_numpy(plan, n, noise_sigma=200.0)
    b = sb.SameReceiverBuilder.samedec(22050)
    o = Oracle(oracle_cfg_from(b))
    o.process_s16(x)
    want = o.events()
    wb = [e for e in want if e.kind == 3]
    assert len(wb) == 1 and len(wb[0].data) > 1400
    for kernel in (0, 1, 2, 4):
        rx = b.build_batch(1)
        rx.set_option("kernel", kernel)
        rx.submit([x]); rx.sync()
        evs, pay = rx.drain_raw()
        got = sb.SameBatchReceiver.events_from_raw(evs, pay)
        assert len(got) == len(want)
        for g, w, r in zip(got, want, evs):
            assert (g.kind, g.err, g.sample, g.symbol_count) == (w.kind, w.err, w.sample, w.symbol_count)
            if g.kind == 3:
                assert int(r["data_len"]) == len(w.data) and (g.flags & 1) and g.data == w.data[:1024]
            else:
                assert g.data == w.data and g.flags == 0


def test_multi_device_entry_two_shards_equal_one_engine():
    """same_multi_*: one batch over several engines, one host thread each (here both shards on device 0; all visible
    devices too when the box has more).  Results == a single engine, global stream ids, sorted."""
    torch = _torch()
    ns = 75
    recs = [synth.render_numpy(synth.plan_stream(2000 + i, seconds=14.0), 14 * 22050 - 31 * i) for i in range(ns)]
    recs[3] = recs[3][:0]
    b = sb.SameReceiverBuilder.samedec(22050)
    want = b.build_batch(ns).process(recs)
    layouts = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() > 1:
        layouts.append(list(range(torch.cuda.device_count())))
    for devices in layouts:
        rx = b.build_multi(ns, devices)
        sh = rx.shards()
        assert sh[0][1] == 0 and sum(c for _, _, c in sh) == ns and all(sh[i][1] + sh[i][2] == sh[i + 1][1] for i in range(len(sh) - 1))
        half = [r[: len(r) // 2] for r in recs]
        rest = [r[len(r) // 2:] for r in recs]
        got = rx.process(half)
        for s_, e in enumerate(rx.process(rest)):
            got[s_].extend(e)
        for s_ in range(ns):
            assert_events_equal(got[s_], want[s_], f"devices {devices} stream {s_}")
        assert np.array_equal(rx.input_sample_counters(), np.array([len(r) for r in recs], np.uint64))
        msgs = list(b.build_multi(ns, devices).iter_messages_batched(recs))
        assert msgs == [(s_, e.message_ok()) for s_ in range(ns) for e in want[s_] if e.message_ok() is not None]
        rx.reset()
        assert not rx.input_sample_counters().any()


def test_kernel_policy_crossovers_are_pinned():
    """The engine picks its kernel from the batch size (same_engine.cu, measured table in profiles/README.md); a change
    of the thresholds must be a deliberate one."""
    _torch()
    b = sb.SameReceiverBuilder.samedec(22050)
    for ns, want in POLICY_TABLE:
        rx = b.build_batch(ns)
        assert rx.get_option("kernel_selected") == want, (ns, rx.get_option("kernel_selected"), want)
        del rx
    assert sb.SameReceiverBuilder.samedec(44100).build_batch(64).get_option("kernel_selected") == 1


# (streams, kernel): 3 pipelined up to one 32-stream block per SM (148 SMs), 4 three-warp up to four, 2 single-warp up to
# eight, 6 look-ahead single-warp beyond
POLICY_TABLE = [(1, 3), (4096, 3), (4736, 3), (4737, 4), (8192, 4), (18944, 4), (18945, 2), (32768, 2), (37888, 2), (37889, 6),
                (65536, 6)]
