"""Multi-GPU plumbing of the encode path: streams are independent (EncoderSession shares nothing, SRC:237-258), so the
batch is partitioned by stream across ranks — one process per GPU, no data-path collective.  torch.distributed is used
only for the barrier and for reducing the timing / byte counters (NCCL on GPUs, gloo in the CPU tests)."""


def shard_range(n_total, rank, world):
    """Contiguous block of streams owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stream_params(i):
    """BASELINE C4 recipe for global stream index i: (f_left, f_right, seed)."""
    fl = 110.0 * 2.0 ** ((i % 48) / 12.0)
    return fl, fl * 1.26, 1000 + i


def _reduce(value, op_name, dist, device):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(value, dist=None, device="cpu"):
    return _reduce(value, "MAX", dist, device)


def sum_over_ranks(value, dist=None, device="cpu"):
    return _reduce(value, "SUM", dist, device)
