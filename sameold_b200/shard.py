"""Multi-GPU sharding of independent streams: contiguous shards, one process + one engine per GPU,
NO collective on the data path.  The only communication is the host-side gather of decoded events/messages
(torch.distributed object gather over whatever backend the job uses: NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_bounds(lengths: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) stream ranges per rank, balanced by total samples (ragged lengths), every rank's range
    possibly empty when there are fewer streams than ranks.  Deterministic and identical on every rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = int(lengths.size)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    total = int(lengths.sum())
    if n == 0:
        return [(0, 0)] * world_size
    if total == 0:  # no samples at all: balance by count
        cuts = [round(i * n / world_size) for i in range(world_size + 1)]
    else:
        cum = np.concatenate([[0], np.cumsum(lengths)])
        cuts = [0]
        for r in range(1, world_size):
            target = total * r / world_size
            # first boundary whose cumulative sample count reaches the target, never moving backwards
            c = int(np.searchsorted(cum, target, side="left"))
            # choose the closer of the two neighbouring boundaries
            if c > 0 and abs(cum[c - 1] - target) <= abs(cum[min(c, n)] - target):
                c -= 1
            cuts.append(min(max(c, cuts[-1]), n))
        cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world_size)]


def local_shard(items: Sequence, lengths: Sequence[int], rank: int, world_size: int):
    a, b = shard_bounds(lengths, world_size)[rank]
    return list(items[a:b]), a


def gather_per_stream(local_results: List, first_stream: int, n_streams: int, group=None) -> List:
    """All ranks contribute `local_results` (one entry per local stream, starting at global index `first_stream`);
    every rank returns the full per-stream list in global order.  Host-side object gather only."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = [None] * n_streams
        out[first_stream:first_stream + len(local_results)] = local_results
        return out
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (first_stream, local_results), group=group)
    out = [None] * n_streams
    for start, res in parts:
        out[start:start + len(res)] = res
    return out


def decode_sharded(recordings: Sequence[np.ndarray], decode_local: Callable[[List[np.ndarray]], List], rank: int,
                   world_size: int, group=None) -> List:
    """Shard `recordings` over the ranks, run `decode_local` (e.g. SameBatchReceiver.decode_samedec on this rank's GPU)
    on the local shard, and gather the per-stream results on every rank."""
    lengths = [len(r) for r in recordings]
    mine, first = local_shard(recordings, lengths, rank, world_size)
    local = decode_local(mine) if mine else []
    if len(local) != len(mine):
        raise RuntimeError("decode_local must return one result per local stream")
    return gather_per_stream(local, first, len(recordings), group)
