"""Multi-GPU plumbing of the one exchange step on the path.

Scene reference points are independent units: rank g of G votes on reference slots g, g+G, g+2G, ...
(interleaved, so that density differences across the scene balance out) against a replicated model
table and scene, then ONE all-gather of fixed-size 64-byte hypothesis records precedes clustering,
which every rank runs redundantly (it is deterministic, so all ranks hold the same poses).

The functions work on CPU tensors (gloo, tests) and CUDA tensors (nccl, bench.py) alike.
"""
from __future__ import annotations

RECORD_FLOATS = 16  # one b200ppf_hypothesis = 64 bytes = 16 float32 words


def shard(n_ref: int, rank: int, world: int):
    """(first slot, slot step, slot count) of one rank."""
    count = (n_ref - rank + world - 1) // world if rank < n_ref else 0
    return rank, world, count


def chunk_size(n_ref: int, world: int) -> int:
    """records every rank contributes to the all-gather (the last slots of some ranks are padding)"""
    return (n_ref + world - 1) // world


def all_gather_hypotheses(local, n_ref: int, world: int, dist=None):
    """local: (chunk, 16) float32 tensor holding this rank's records in slot order (padding rows at
    the end).  Returns a contiguous (n_ref_padded, 16) tensor whose first n_ref rows are the records
    of reference slots 0..n_ref-1 in order — exactly what a single rank would have produced."""
    import torch
    chunk = local.shape[0]
    if world == 1:
        return local
    gathered = torch.empty((world * chunk, RECORD_FLOATS), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local.contiguous())
    # rank-major [world][chunk] -> slot order k = c * world + g ; padding ends up at the tail
    return gathered.view(world, chunk, RECORD_FLOATS).transpose(0, 1).contiguous().view(world * chunk, RECORD_FLOATS)
