"""N > 1 path on CPU: world_size 2 over gloo.  Each rank encodes its shard of a small batch (with the CPU oracle
standing in for the device, there is no GPU here), then the byte counters / timings are reduced exactly as bench.py does."""
import hashlib
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
N_STREAMS = 5


def _stream(i, sh):
    import signals
    fl, fr, seed = sh.stream_params(i)
    return signals.sine_noise(0.2, f_left=fl, f_right=fr, seed=seed)


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("swift-mp3_b200.sharding")
    import oracle_binding as orc
    lo, hi = sh.shard_range(N_STREAMS, rank, world)
    digests, nbytes = {}, 0
    for i in range(lo, hi):
        out, _ = orc.encode_all(_stream(i, sh))
        digests[i] = hashlib.sha256(out).hexdigest(); nbytes += len(out)
    dist.barrier()
    total = sh.sum_over_ranks(nbytes, dist)
    slowest = sh.max_over_ranks(1.0 + rank, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    if rank == 0:
        merged = {}
        for g in gathered:
            assert not set(g) & set(merged), "shards overlap"
            merged.update(g)
        q.put((merged, total, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges(mp3):
    sh = importlib.import_module("swift-mp3_b200.sharding")
    for n in (0, 1, 5, 8, 4096):
        for w in (1, 2, 3, 4, 8):
            r = [sh.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert sh.stream_params(0) == (110.0, 110.0 * 1.26, 1000) and sh.stream_params(48)[0] == 110.0
    assert sh.max_over_ranks(3.5) == 3.5 and sh.sum_over_ranks(2) == 2.0


def test_two_ranks_gloo(orc):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    merged, total, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    sh = importlib.import_module("swift-mp3_b200.sharding")
    want, nbytes = {}, 0
    for i in range(N_STREAMS):
        out, _ = orc.encode_all(_stream(i, sh))
        want[i] = hashlib.sha256(out).hexdigest(); nbytes += len(out)
    assert merged == want and total == nbytes and slowest == 2.0
