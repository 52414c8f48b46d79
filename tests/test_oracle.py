"""CPU tests: the oracle against the reference's golden files and unit-test KATs."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import GOLDEN_DIR, Oracle, build_oracle, load_golden_recording
from oracle.pyoracle import HERE as ORACLE_DIR, default_config

NAMES = ["long_message", "npt", "two_and_two"]


def expected_messages(name):
    # sample/<name>.22050.s16le.txt = samedec stdout; "+OK" lines come from the child script, not the decoder
    with open(os.path.join(GOLDEN_DIR, f"{name}.22050.s16le.txt")) as f:
        return [l.rstrip("\n") for l in f if not l.startswith("+OK")]


def test_selftest_kats():
    """oracle/selftest.cpp mirrors the reference's #[test] known answers (dcblock, agc, filter, waveform, demod,
    symsync, codesquelch, equalize, framing, combiner, assembler, message, receiver)."""
    build_oracle()
    out = subprocess.run([os.path.join(ORACLE_DIR, "_build", "selftest")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("OK ")
    assert int(out.stdout.split()[1]) > 500


@pytest.mark.parametrize("name", NAMES)
def test_golden_recordings(name):
    """BASELINE config 1: samedec semantics on sample/*.bin reproduce sample/*.txt (sample/test.sh:21-57)."""
    o = Oracle.samedec(22050)
    o.process_s16(load_golden_recording(name))
    o.flush_samedec()
    assert o.messages() == expected_messages(name)


def test_golden_event_trace_is_stable():
    with open(os.path.join(GOLDEN_DIR, "oracle_events.json")) as f:
        gold = json.load(f)
    for name in NAMES:
        o = Oracle.samedec(22050)
        o.process_s16(load_golden_recording(name))
        o.flush_samedec()
        assert [e.to_json() for e in o.events()] == gold[name]


def test_long_message_needs_flush():
    """the long_message header is only released 9k samples into the first EOF flush."""
    s = load_golden_recording("long_message")
    o = Oracle.samedec()
    o.process_s16(s)
    assert o.messages() == []
    o.flush_samedec()
    assert len(o.messages()) == 1


def test_chunked_equals_whole():
    """receiver state persists across iter_events calls (receiver.rs:233-274): any chunking gives the same events."""
    s = load_golden_recording("npt")
    a = Oracle.samedec()
    a.process_s16(s)
    b = Oracle.samedec()
    rng = np.random.default_rng(1)
    i = 0
    while i < len(s):
        n = int(rng.integers(1, 5000))
        b.process_s16(s[i:i + n])
        i += n
    assert [e.key() for e in a.events()] == [e.key() for e in b.events()]


def test_derived_constants():
    """receiver.rs:502-560 at 22050 Hz, samedec config (values quoted in SURVEY §3.4)."""
    d = Oracle.derived(default_config(22050, True))
    assert d["dc_len"] == 16 and d["ntaps"] == 42
    assert abs(d["sps"] - 42.336269) < 1e-5
    assert abs(d["agc_bw"] - 1.9200121e-5) < 1e-11
    assert abs(d["agc_gain0"] - 1.0 / 32767.0) < 1e-10
    assert abs(d["samples_per_ted"] - 21.168135) < 1e-5
    assert abs(d["period_min"] - 20.744772) < 1e-5 and abs(d["period_max"] - 21.591497) < 1e-5
    assert abs(d["alpha_unlocked"] - 0.79212046) < 1e-6 and abs(d["beta_unlocked"] - 0.29600334) < 1e-6
    assert abs(d["alpha_locked"] - 0.46651193) < 1e-6 and abs(d["beta_locked"] - 0.07268274) < 1e-6
    # taps: |h| = 2/42, newest sample pairs with tap 0 whose phase is 2*pi*f/fs*41
    assert np.allclose(np.hypot(d["mark"][:, 0], d["mark"][:, 1]), 2.0 / 42.0, atol=1e-7)
    assert d["mark"][41, 0] == np.float32(2.0) / np.float32(42.0) and d["mark"][41, 1] == 0.0


def test_invalid_config_rejected():
    cfg = default_config()
    cfg.dc_blocker_len = 0.0  # MovingAverage::new asserts len > 0 (dcblock.rs:74)
    with pytest.raises(ValueError):
        Oracle(cfg)


def test_empty_and_tiny_inputs():
    o = Oracle.samedec()
    o.process_s16(np.zeros(0, np.int16))
    assert o.events() == [] and o.input_sample_counter == 0
    o.process_s16(np.array([123], np.int16))
    assert o.events() == [] and o.input_sample_counter == 1


def test_soft_trace_symbol_rate():
    s = load_golden_recording("npt")
    o = Oracle.samedec()
    o.enable_trace()
    o.process_s16(s)
    t = o.soft_trace()
    # ~520.83 symbols/s
    assert abs(len(t) / (len(s) / 22050.0) - 520.83) < 2.0
    assert np.all(np.abs(t["sym"]) <= 1.0)


def test_batch_events_export_equals_the_single_stream_api():
    """oracle_decode_batch_events (what the full-size GPU parity tests compare whole arrays against) must hand back
    exactly what the single-receiver API reports, stream by stream, with and without samedec's EOF flush, for ragged
    lengths and any thread count."""
    from oracle import decode_batch_events
    from oracle.pyoracle import default_config
    recs = [load_golden_recording(n) for n in ("npt", "two_and_two")]
    n = max(len(r) for r in recs)
    a = np.zeros((3, n), np.int16)
    a[0, :len(recs[0])] = recs[0]
    a[1, :len(recs[1])] = recs[1]
    a[2, :60000] = recs[0][:60000]
    lengths = [len(recs[0]), len(recs[1]), 60000]
    for flush in (False, True):
        for threads in (1, 3):
            ev, pay, _ = decode_batch_events(default_config(), a, threads, lengths=lengths, flush=flush)
            assert np.all(np.diff(ev["stream"].astype(np.int64)) >= 0)
            for s in range(3):
                o = Oracle.samedec()
                o.process_s16(a[s, :lengths[s]])
                if flush:
                    o.flush_samedec()
                want = o.events()
                got = ev[ev["stream"] == s]
                assert got.size == len(want)
                assert list(got["seq"]) == list(range(len(want)))
                for g, w in zip(got, want):
                    assert (g["kind"], g["err"], g["sample"], g["symbol_count"], g["parity_errors"], g["voting_bytes"]) == \
                        (w.kind, w.err, w.sample, w.symbol_count, w.parity_errors, w.voting_bytes)
                    assert bytes(pay[g["data_offset"]:g["data_offset"] + g["data_len"]]) == w.data


def test_cpu_corpus_generator_is_deterministic_and_decodable():
    """oracle/synth_cpu.hpp (the reference arm's corpus): same plans -> same samples for any thread count; the streams
    carry their planned headers at 10 dB SNR."""
    from oracle import decode_batch_events, synth_cpu
    from oracle.pyoracle import default_config
    from sameold_b200 import synth
    plans = synth.plan_corpus(6, 22050, 45.0, first_stream=31)
    x1 = synth_cpu(plans, 45 * 22050, 22050, 1)
    x4 = synth_cpu(plans, 45 * 22050, 22050, 4)
    assert np.array_equal(x1, x4) and x1.dtype == np.int16 and x1.shape == (6, 45 * 22050)
    assert 3000 < x1[:, :20000].std() < 4500            # AWGN sigma 3663 before the first burst
    ev, pay, _ = decode_batch_events(default_config(), x1, 2)
    som = ev[ev["kind"] == 18]
    texts = {int(e["stream"]): bytes(pay[e["data_offset"]:e["data_offset"] + e["data_len"]]).decode() for e in som}
    assert sum(texts.get(i) == plans[i].header for i in range(6)) >= 5
