"""CPU: the multi-GPU host logic — round-robin sharding of the pair list and the gather of result
records — with world_size 2 and 3 over gloo.  Local compute is stood in by the oracle (tests may
use it); the point is that record k of the global batch comes back at slot k on every rank."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from dpg_slam_b200 import sharded
from dpg_slam_b200._abi import RESULT_DTYPE


def test_shard_indices_partition():
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            parts = [sharded.shard_indices(n, r, w) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            for r in range(w):
                assert len(parts[r]) == sharded.shard_len(n, r, w) <= sharded.padded_len(n, w)
                assert np.all(parts[r] % w == r)


def test_interleave_restores_global_order():
    for n, w in ((10, 3), (8, 2), (5, 8), (0, 2)):
        m = sharded.padded_len(n, w)
        g = np.zeros((w, m), RESULT_DTYPE)
        for r in range(w):
            idx = sharded.shard_indices(n, r, w)
            g[r, :len(idx)]["iterations"] = idx
        out = sharded.interleave(g, n, w)
        assert out["iterations"].tolist() == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dpg_slam_b200 import synth
        from dpg_slam_b200._abi import COV_CENSI_CORR, Params
        from oracle import oracle_py as O
        wl = synth.config_corridor(n_pairs=n_pairs, n_beams=181, seed=9)
        pts, off = O.clouds_from_ranges(wl.ranges, wl.scanner)
        p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR)
        idx = sharded.shard_indices(n_pairs, rank, world)
        local, _ = O.run_batch(pts, off, wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx], p, fast=1)
        allrec = sharded.gather_records(local, n_pairs, rank, world)
        full, _ = O.run_batch(pts, off, wl.src_idx, wl.tgt_idx, wl.guess, p, fast=1)
        q.put((rank, allrec.tobytes() == full.tobytes(), len(allrec)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pairs", [(2, 9), (3, 10)])
def test_gloo_gather_equals_single_rank(world, n_pairs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in got) == list(range(world))
    assert all(ok and n == n_pairs for _, ok, n in got)
