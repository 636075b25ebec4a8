"""
Multi-GPU layout of the gated update path: independent video streams, one stream group per rank.

Every stream owns its gate / buffer / accumulator state and never reads another stream's (the reference
resets per video, scripts/time/vitdet_vid.py:26), so ranks shard streams and the hot path has NO
collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to agree on timing
(max over ranks) and to collect per-stream outputs after the timed region.
"""

import torch


def partition_streams(total_streams, world_size, rank):
    """Contiguous, balanced slice of stream ids for `rank` (earlier ranks take the remainder)."""
    if total_streams < 0 or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad partition request: streams={total_streams} world={world_size} rank={rank}")
    base, extra = divmod(total_streams, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def stream_seed(base_seed, stream_id):
    """Synthetic-video seed of a stream: a function of the GLOBAL stream id, not of the rank layout."""
    return base_seed + 7919 * stream_id


def max_over_ranks(value, dist=None, device="cpu"):
    """Timing agreement: every rank gets the slowest rank's elapsed time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_stream_outputs(local, stream_ids, total_streams, dist=None):
    """
    Collects per-stream outputs on every rank, ordered by global stream id.
    local: (len(stream_ids), ...) tensor of this rank's streams.  Ranks may hold different stream counts.
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [len(partition_streams(total_streams, world, r)) for r in range(world)]
    width = max(counts)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    assert out.shape[0] == total_streams and len(stream_ids) == counts[dist.get_rank()]
    return out
