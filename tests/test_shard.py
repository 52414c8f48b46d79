"""CPU tests of the N>1 path: contiguous sharding balanced by samples, and a world_size-2 gloo run that shards the
golden recordings over two processes, decodes each shard and gathers the messages on every rank (no data-path
collective).  The per-rank decoder here is the CPU oracle (test infrastructure) — on GPUs it is
SameBatchReceiver.decode_samedec."""
import os
import subprocess
import sys

import numpy as np
import pytest

from sameold_b200.shard import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    for n, w in [(0, 4), (1, 4), (3, 2), (7, 3), (4096, 8), (65536, 8), (100, 7)]:
        lengths = rng.integers(0, 1_400_000, n)
        b = shard_bounds(lengths, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1)) and all(s <= e for s, e in b)
        if n >= 8 * w:
            per = [int(lengths[s:e].sum()) for s, e in b]
            assert max(per) - min(per) <= 2 * int(lengths.max())   # balanced to within one stream each side
    # equal lengths split evenly
    assert shard_bounds([10] * 8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert shard_bounds([5, 5, 5], 1) == [(0, 3)]
    with pytest.raises(ValueError):
        shard_bounds([1], 0)


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SAME_ROOT"])
import torch.distributed as dist
from oracle import Oracle, load_golden_recording
from sameold_b200.shard import decode_sharded

def decode_local(recs):
    out = []
    for r in recs:
        o = Oracle.samedec(22050); o.process_s16(r); o.flush_samedec(); out.append(o.messages())
    return out

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
names = ["long_message", "npt", "two_and_two", "npt", "two_and_two"]
recs = [load_golden_recording(n) for n in names]
res = decode_sharded(recs, decode_local, rank, world)
expected = {}
for n in set(names):
    with open(os.path.join(os.environ["SAME_ROOT"], "tests", "golden", f"{n}.22050.s16le.txt")) as f:
        expected[n] = [l.rstrip("\n") for l in f if not l.startswith("+OK")]
assert res == [expected[n] for n in names], (rank, res)
dist.barrier()
if rank == 0:
    print("SHARD_OK", world)
dist.destroy_process_group()
'''


def test_two_rank_gloo_decode_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SAME_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARD_OK 2" in out.stdout


GPU_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SAME_ROOT"])
import torch, torch.distributed as dist
import sameold_b200 as sb
from oracle import load_golden_recording
from sameold_b200.shard import decode_sharded

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
device = rank % torch.cuda.device_count()        # one engine per rank; ranks share a GPU when the box has fewer

def decode_local(recs):
    rx = sb.SameReceiverBuilder.samedec(22050).build_batch(len(recs), device=device)
    return rx.decode_samedec(recs)

names = ["long_message", "npt", "two_and_two", "npt", "two_and_two"]
recs = [load_golden_recording(n) for n in names]
res = decode_sharded(recs, decode_local, rank, world)
expected = {}
for n in set(names):
    with open(os.path.join(os.environ["SAME_ROOT"], "tests", "golden", f"{n}.22050.s16le.txt")) as f:
        expected[n] = [l.rstrip("\n") for l in f if not l.startswith("+OK")]
assert res == [expected[n] for n in names], (rank, res)
dist.barrier()
if rank == 0:
    print("GPU_SHARD_OK", world)
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_two_rank_decode_sharded_with_the_gpu_engine(tmp_path):
    """decode_sharded over two processes, each with its own CUDA engine (SameBatchReceiver.decode_samedec): the
    multi-process host path bench.py's torchrun launch uses, with the product decoder instead of the oracle."""
    script = tmp_path / "gpu_worker.py"
    script.write_text(GPU_WORKER)
    env = dict(os.environ, SAME_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GPU_SHARD_OK 2" in out.stdout
