"""The C++ host layer (include/same_receiver.hpp) over the C ABI: builds everywhere; without a GPU it must refuse with
SAME_ERR_NO_DEVICE; on the GPU box it decodes the golden recordings like the reference's receiver tests."""
import gzip
import os
import subprocess

import pytest

import sameold_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["long_message", "npt", "two_and_two"]


def _build(tmp_path):
    lib = sb.build_native()
    exe = str(tmp_path / "test_receiver")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "test_receiver.cpp"), "-L" + os.path.dirname(lib), "-lsame_b200",
           "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    bins = []
    for n in NAMES:
        p = tmp_path / f"{n}.bin"
        with gzip.open(os.path.join(ROOT, "tests", "golden", f"{n}.22050.s16le.bin.gz"), "rb") as f:
            p.write_bytes(f.read())
        bins.append(str(p))
    return exe, bins


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_cpp_layer_builds_and_refuses_without_gpu(tmp_path):
    exe, bins = _build(tmp_path)
    r = subprocess.run([exe] + bins, capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)       # SAME_ERR_NO_DEVICE, surfaced as same::EngineError
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_layer_decodes_golden_recordings(tmp_path):
    exe, bins = _build(tmp_path)
    r = subprocess.run([exe] + bins, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "CPP_OK" in r.stdout
