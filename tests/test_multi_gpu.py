"""GPU, needs >= 2 devices (skipped on a 1-GPU box): the sharded path over NCCL gives, on every rank, the
same records bit for bit as one GPU running the whole batch."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from dpg_slam_b200 import sharded, synth
        from dpg_slam_b200._abi import COV_CENSI_CORR, Params
        from dpg_slam_b200.scanmatch import ScanMatcher
        wl = synth.config_loop_closure(n_pairs=n_pairs, n_scans=200, seed=31)
        p = Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR)
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream), ScanMatcher(rank) as sm:
            sm.set_stream(stream.cuda_stream)
            sm.upload_ranges(wl.ranges, wl.scanner)                 # scan store replicated on every GPU
            sh = sharded.ShardedScanMatcher(sm, rank, world, None, dev)
            sh.set_pairs(wl.src_idx, wl.tgt_idx, wl.guess)          # pair k -> rank k % world
            sh.run(p)
            allrec = sh.gather()                                    # NCCL all-gather of the records
            # the same gather fused into the kernel epilogue: peer stores into every rank's buffer, no collective
            sh.attach_fused_gather(n_pairs)
            sh.run(p)
            fused = sh.fused_records()
            sh.run(p)                                               # buffers are reusable batch after batch
            fused2 = sh.fused_records()
            sh.detach_fused_gather()
            assert fused.tobytes() == allrec.tobytes() and fused2.tobytes() == allrec.tobytes()
            # root-only gather: only rank 0's buffer receives the batch (what one host-side pose-graph update needs)
            sh.attach_fused_gather(n_pairs, root_only=True)
            sh.run(p)
            root = sh.fused_records()
            sh.detach_fused_gather()
            assert (root is None) == (rank != 0)
            if rank == 0:
                assert root.tobytes() == allrec.tobytes()
            # device-resident caller, sharded: every rank enumerates the list of one reoptimize() on its device and keeps
            # its round-robin shard; records gathered into rank 0's buffer == rank 0 alone running the whole list
            from dpg_slam_b200._abi import ENUM_REOPTIMIZE
            ms = synth.config_multisession(n_sessions=2, scans_per_session=100, n_beams=541, seed=9, size=30.0, n_boxes=30)
            pd = Params.defaults(cov_mode=COV_CENSI_CORR)
            sm.upload_ranges(ms.ranges, ms.scanner)
            sm.set_nodes(ms.poses_est, ms.passes)
            total, local = sm.enumerate_pairs_device(ENUM_REOPTIMIZE, 5.0, 2.0, rank, world)
            sh.n_pairs = total
            sh.attach_fused_gather(total, root_only=True)
            sm.run(pd)
            enum_all = sh.fused_records(total)
            sh.detach_fused_gather()
            single = None
            ok_enum = True
            if rank == 0:
                t1, l1 = sm.enumerate_pairs_device(ENUM_REOPTIMIZE, 5.0, 2.0, 0, 1)
                sm.run(pd)
                ok_enum = t1 == total and sm.fetch_results().tobytes() == enum_all.tobytes()
                sm.upload_ranges(wl.ranges, wl.scanner)
                sm.set_pairs(wl.src_idx, wl.tgt_idx, wl.guess)
                sm.run(p)
                single = sm.fetch_results()
        ok = ok_enum and (True if single is None else allrec.tobytes() == single.tobytes())
        q.put((rank, ok, len(allrec), int(allrec["iterations"].sum())))
    finally:
        dist.destroy_process_group()


def test_sharded_nccl_equals_single_gpu():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    n_pairs = 1001                                                  # not a multiple of the world size
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ok and n == n_pairs for _, ok, n, _ in got)
    assert len({s for _, _, _, s in got}) == 1                      # every rank holds the same gathered batch
