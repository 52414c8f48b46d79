"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares, and fails
loudly (never falls back) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sameold_b200 as sb
from sameold_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(same_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(sb.build_native())
    names = _declared_functions("same_engine.h") + _declared_functions("same_synth.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported by libsame_b200.so"
    # and the Python binding knows all of them
    for n in names:
        assert n in _lib.API, f"{n} missing from sameold_b200/_lib.py"


def test_abi_version_and_struct_sizes():
    lib = _lib.load()
    assert lib.same_abi_version() == 2
    assert C.sizeof(_lib.SameConfig) == 19 * 4
    assert C.sizeof(_lib.SameEvent) == 48
    assert C.sizeof(_lib.SameSoftSymbol) == 16


def test_config_defaults_match_builder_rs():
    """SameReceiverBuilder::new (builder.rs:50-67) and samedec's overrides (main.rs:29-37)."""
    b = sb.SameReceiverBuilder(22050)
    assert b.input_rate() == 22050
    assert b.dc_blocker_length() == np.float32(0.38)
    assert b.agc_bandwidth() == np.float32(0.01)
    assert b.agc_gain_limits() == (0.0, np.float32(1.0e6))
    assert b.timing_bandwidth() == (0.125, np.float32(0.05))
    assert b.timing_max_deviation() == np.float32(0.01)
    assert b.squelch_power() == (np.float32(0.10), np.float32(0.05))
    assert b.squelch_bandwidth() == 0.125
    assert b.preamble_max_errors() == 2
    assert b.frame_prefix_max_errors() == 2 and b.frame_max_invalid() == 5
    eq = b.adaptive_equalizer()
    assert eq.filter_order() == (6, 4) and eq.relaxation() == np.float32(0.05) and eq.regularization() == np.float32(1e-6)
    s = sb.SameReceiverBuilder.samedec(22050)
    assert s.agc_gain_limits() == (np.float32(1.0) / np.float32(32767.0), np.float32(1.0) / np.float32(200.0))


def test_builder_setters_clamp_like_the_reference():
    """builder.rs:95-279, 393-425 (and its tests builder.rs:428-449)."""
    b = sb.SameReceiverBuilder(48000)
    b.with_dc_blocker_length(-1.0)
    assert b.dc_blocker_length() == 0.0
    b.with_agc_bandwidth(2.0)
    assert b.agc_bandwidth() == 1.0
    b.with_timing_bandwidth(0.5, 0.9)       # locked is clamped to <= unlocked
    assert b.timing_bandwidth() == (0.5, 0.5)
    b.with_timing_max_deviation(0.9)
    assert b.timing_max_deviation() == 0.5
    b.with_squelch_power(2.0, 3.0)          # open clamped to 1, close = min(close, open-as-given)
    assert b.squelch_power() == (1.0, 2.0)
    b.with_frame_prefix_max_errors(100)
    assert b.frame_prefix_max_errors() == 7
    e = sb.EqualizerBuilder().with_filter_order(0, 9).with_relaxation(7.0).with_regularization(-1.0)
    assert e.filter_order() == (1, 1) and e.relaxation() == 1.0 and e.regularization() == 0.0
    b.without_adaptive_equalizer()
    assert b.adaptive_equalizer() is None
    b.with_adaptive_equalizer(sb.EqualizerBuilder().with_filter_order(8, 4))
    assert b.adaptive_equalizer().filter_order() == (8, 4)
    # sanitize in C agrees with the Python setters
    cfg = b.config()
    _lib.load().same_config_sanitize(C.byref(cfg))
    assert (cfg.timing_bw_unlocked, cfg.timing_bw_locked) == (0.5, 0.5)


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_is_a_loud_error_not_a_fallback():
    with pytest.raises(sb.SameEngineError) as ei:
        sb.SameReceiverBuilder(22050).build()
    assert ei.value.code == 3  # SAME_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; nothing under sameold_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sameold_b200")):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "same_oracle" not in txt and "liboracle" not in txt and "from oracle" not in txt \
                    and "import oracle" not in txt, f"{f} references the oracle"


def test_rust_ffi_declares_the_whole_header():
    """bindings/rust/src/ffi.rs (source only: no Rust toolchain here) must declare every function of the C header,
    and its #[repr(C)] structs must list the header's fields in the same order."""
    import re
    hdr = open(os.path.join(ROOT, "include", "same_engine.h")).read()
    ffi = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    c_fns = set(re.findall(r"^[A-Za-z_][A-Za-z0-9_ \*]*?\b(same_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    rust_fns = set(re.findall(r"pub fn (same_[a-z0-9_]+)", ffi))
    assert c_fns and c_fns == rust_fns, (sorted(c_fns - rust_fns), sorted(rust_fns - c_fns))

    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                out += [f.strip() for f in decl.split(None, 1)[1].split(",")]
        return out

    def rust_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, ffi, flags=re.S).group(1)
        return re.findall(r"pub ([a-z0-9_]+):", body)

    for name in ("same_config", "same_event", "same_soft_symbol", "same_derived"):
        assert c_fields(name) == rust_fields(name), name


def test_library_reads_no_environment_variables():
    """a host application's environment must not be able to override the measured kernel policy; the only
    override is same_engine_set_option."""
    for f in os.listdir(os.path.join(ROOT, "sameold_b200", "csrc")):
        txt = open(os.path.join(ROOT, "sameold_b200", "csrc", f), errors="ignore").read()
        assert "getenv" not in txt, f


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_multi_device_entry_is_loud_without_gpu():
    with pytest.raises(sb.SameEngineError) as ei:
        sb.SameReceiverBuilder(22050).build_multi(64, [0, 0])
    assert ei.value.code == 3
