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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_near_gpu(index):
    """One process per GPU: run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned
    buffer is allocated, so that the staging memory is local to the GPU's PCIe root (first-touch placement).  With eight
    ranks pulling 55 GB/s each, buffers that all sit on one socket are bound by the inter-socket link instead of PCIe.
    Returns a short description for the bench record; does nothing where the topology is not exposed."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + addr
        node = int(open(base + "/numa_node").read().strip())
        cpus = _parse_cpulist(open(base + "/local_cpulist").read()) & set(os.sched_getaffinity(0))
        if node < 0 or not cpus:
            return {"gpu": addr, "numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"gpu": addr, "numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:   # no sysfs topology (virtualised box), old torch, ...
        return {"bound": False, "why": str(e)[:80]}
