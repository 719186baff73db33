"""N>1 host logic on CPU: two gloo processes shard pairs round-robin, run a stand-in registration per pair with no
data-path collective, and reduce the job statistics in one all_gather (rdmnet_b200/sharding.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rdmnet_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _load(pid):
    rng = np.random.default_rng(100 + pid)
    return rng.standard_normal((50, 3)).astype(np.float32), rng.standard_normal((40, 3)).astype(np.float32)


def _register(ref, src):  # stand-in for PairRegistrar.register: any pure function of the pair
    return float(ref.sum() - src.sum())


def _worker(rank, world, port, num_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = sharding.run_sharded(num_pairs, rank, world, _register, _load)
        stats = sharding.reduce_job_stats(len(res), 0.5 + rank, extra={"checksum": sum(res.values())})
        q.put((rank, sorted(res), stats))
    finally:
        dist.destroy_process_group()


def test_shard_ids_partition():
    for n in (0, 1, 7, 8, 256):
        for w in (1, 2, 3, 8):
            got = sorted(i for r in range(w) for i in sharding.shard_pair_ids(n, r, w))
            assert got == list(range(n))
            padded = [sharding.shard_pair_ids(n, r, w, pad=True) for r in range(w)]
            assert len({len(x) for x in padded}) == 1
            if n:
                assert set(i for x in padded for i in x) == set(range(n))


def test_two_rank_gloo_job():
    world, num_pairs = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5]
    want = sum(_register(*_load(i)) for i in range(num_pairs))
    for _, _, st in out:  # every rank sees the same whole-job numbers
        assert st["pairs"] == num_pairs and st["elapsed_s"] == 1.5
        assert abs(st["pairs_per_s"] - num_pairs / 1.5) < 1e-12
        assert abs(st["checksum"] - want) < 1e-6
        assert st["per_rank_pairs"] == [4.0, 3.0]


def test_single_process_stats():
    st = sharding.reduce_job_stats(10, 2.0)
    assert st["pairs_per_s"] == 5.0 and st["per_rank_pairs"] == [10.0]
