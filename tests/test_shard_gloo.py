"""Multi-process (world_size 2, gloo, CPU) test of the sharding / gather host
logic with a stub scorer standing in for the per-rank GPU engine."""
import os
import socket

import numpy as np
import pytest

from nele_gan_b200 import shard


def test_partition_is_a_balanced_permutation():
    rng = np.random.default_rng(0)
    lens = rng.integers(16000 * 3, 16000 * 10, size=1001)
    for world in (1, 2, 3, 8):
        parts = shard.partition(lens, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(lens)))
        loads = [lens[p].sum() for p in parts]
        assert max(loads) - min(loads) <= lens.max()          # length-sorted deal balances ragged batches
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


class _StubResult:
    def __init__(self, refs, degs):
        n = len(refs)
        self.scores = np.array([[r.sum(), d.sum(), len(r)] for r, d in zip(refs, degs)], dtype=np.float64).reshape(n, 3)
        self.haspi_raw = np.array([r[:10] for r in refs], dtype=np.float64).reshape(n, 10)
        self.status = np.arange(n, dtype=np.int32) * 0 + 0x020100


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    refs = [rng.standard_normal(int(n)).astype(np.float32) for n in rng.integers(20, 60, size=11)]
    degs = [r * 2 for r in refs]
    full = shard.score_sharded(lambda a, b: _StubResult(a, b), refs, degs)
    want = shard.pack_records(_StubResult(refs, degs))
    q.put((rank, bool(np.array_equal(full, want))))
    dist.destroy_process_group()


def test_two_rank_gather_restores_input_order():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
