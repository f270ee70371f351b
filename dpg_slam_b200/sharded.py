"""Multi-GPU form of the batch: scan pairs sharded round-robin, records gathered back.

Scan pairs are independent until the pose-graph update (reference src/dpg_slam/dpg_slam.cc:119,313
run optimizeGraph only after every runIcp call of the step), so pair ``k`` goes to rank
``k % world`` with the scan store replicated on every GPU and NO data-path collective.  The only
exchange is the final gather of the fixed-size result records (112 B each) that the host-side
pose-graph update consumes: one ``all_gather`` over NCCL (NVLink/NVSwitch) on the device buffer the
kernel wrote, or over gloo on host arrays in the CPU tests.

One process per GPU (``torch.distributed``); torch is plumbing here (process group, device
tensors), the arithmetic is the C-ABI library's.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from ._abi import RESULT_DTYPE

RECORD_BYTES = RESULT_DTYPE.itemsize


def shard_indices(n_pairs: int, rank: int, world: int) -> np.ndarray:
    """Global pair indices owned by ``rank``: k with k % world == rank, ascending."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_pairs, world, dtype=np.int64)


def shard_len(n_pairs: int, rank: int, world: int) -> int:
    return (n_pairs - rank + world - 1) // world if n_pairs > rank else 0


def padded_len(n_pairs: int, world: int) -> int:
    """Records per rank in the gathered buffer (every rank contributes the same count)."""
    return (n_pairs + world - 1) // world


def interleave(gathered: np.ndarray, n_pairs: int, world: int) -> np.ndarray:
    """``gathered`` is (world, padded_len) records in rank-major order -> global pair order.
    Record k sits at [k % world, k // world]; the transpose view makes that a plain reshape."""
    g = np.asarray(gathered).reshape(world, padded_len(n_pairs, world))
    return np.ascontiguousarray(g.T).reshape(-1)[:n_pairs]


class _DevMem:
    """Minimal ``__cuda_array_interface__`` holder for a raw device pointer (no ownership)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


def device_bytes_tensor(ptr: int, nbytes: int, device):
    import torch
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


def gather_records(local: np.ndarray, n_pairs: int, rank: int, world: int, group=None,
                   device_ptr: Optional[int] = None, device=None) -> np.ndarray:
    """All-gather the per-rank record arrays and return all ``n_pairs`` records in global order on
    every rank.  ``local`` holds this rank's ``shard_len`` records (host); when ``device_ptr`` is
    given (the address from ``dpgicp_results_device_ptr``) the gather runs device-to-device over
    NCCL without a host bounce and ``local`` may be None."""
    import torch
    import torch.distributed as dist

    m = padded_len(n_pairs, world)
    mine = shard_len(n_pairs, rank, world)
    if device_ptr is not None:
        send = torch.zeros(m * RECORD_BYTES, dtype=torch.uint8, device=device)
        if mine:
            send[:mine * RECORD_BYTES].copy_(device_bytes_tensor(device_ptr, mine * RECORD_BYTES, device))
        recv = torch.empty(world * m * RECORD_BYTES, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(recv, send, group=group)
        host = recv.cpu().numpy()
    else:
        buf = np.zeros(m, RESULT_DTYPE)
        if mine:
            buf[:mine] = local[:mine]
        send = torch.from_numpy(buf.view(np.uint8).copy())
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        host = torch.cat(parts).numpy()
    return interleave(host.view(RESULT_DTYPE), n_pairs, world)


class ShardedScanMatcher:
    """One rank's view of the sharded batch: replicated scan store, round-robin pair shard.

    ``matcher`` is this rank's :class:`dpg_slam_b200.scanmatch.ScanMatcher` (one GPU).  The matcher
    must run on the torch stream the collective is issued on (``matcher.set_stream``) so that the
    gather is ordered after the kernel without a host synchronisation."""

    def __init__(self, matcher, rank: int, world: int, group=None, device=None):
        self.sm, self.rank, self.world, self.group, self.device = matcher, rank, world, group, device
        self.n_pairs = 0
        self._send = self._recv = None
        self._remap = None
        self._fused = 0

    def plan_subset(self, src_idx, tgt_idx, n_scans_total: int) -> np.ndarray:
        """The distinct scans this rank's pairs touch (ascending) — the rows of a per-rank store — and the
        remapping of global scan ids to those rows, remembered for :meth:`set_pairs`."""
        idx = shard_indices(int(len(src_idx)), self.rank, self.world)
        used = np.unique(np.concatenate([np.asarray(src_idx)[idx], np.asarray(tgt_idx)[idx]])).astype(np.int32)
        remap = np.full(int(n_scans_total), -1, np.int32)
        remap[used] = np.arange(used.shape[0], dtype=np.int32)
        self._remap = remap
        return used

    def upload_ranges_for_shard(self, ranges, scanner, src_idx, tgt_idx, n_scans_total=None, n_beams=None):
        """Instead of replicating the whole scan store, upload only the scans this rank's pairs touch.  ``ranges``
        is the full host array (or a raw pointer to page-locked memory, read in place by the kernel)."""
        total = int(n_scans_total) if n_scans_total is not None else int(ranges.shape[0])
        used = self.plan_subset(src_idx, tgt_idx, total)
        self.sm.upload_ranges_subset(ranges, used, scanner, n_scans_total=n_scans_total, n_beams=n_beams)
        return used

    def use_full_store(self):
        """Forget a subset plan: the store holds every scan again (global scan ids)."""
        self._remap = None

    def set_pairs(self, src_idx, tgt_idx, guess):
        """Takes the GLOBAL pair list; keeps this rank's shard resident on its GPU."""
        import torch
        self.n_pairs = int(len(src_idx))
        idx = shard_indices(self.n_pairs, self.rank, self.world)
        g = np.ascontiguousarray(guess, np.float32).reshape(-1, 3)
        s_loc, t_loc = np.asarray(src_idx)[idx], np.asarray(tgt_idx)[idx]
        if getattr(self, "_remap", None) is not None:       # the store holds only this shard's scans
            s_loc, t_loc = self._remap[s_loc], self._remap[t_loc]
        self.sm.set_pairs(s_loc, t_loc, g[idx])
        m = padded_len(self.n_pairs, self.world)
        if self._send is None or self._send.numel() != m * RECORD_BYTES:
            self._send = torch.zeros(m * RECORD_BYTES, dtype=torch.uint8, device=self.device)
            self._recv = torch.empty(self.world * m * RECORD_BYTES, dtype=torch.uint8, device=self.device)

    def run(self, params):
        self.sm.run(params)

    def gather_device(self):
        """NCCL all-gather straight from the kernel's record buffer (device to device, asynchronous on
        the current torch stream).  Returns the (world * padded_len * 112)-byte device tensor in
        rank-major order."""
        import torch.distributed as dist
        ptr, n = self.sm.results_device_ptr()
        mine = shard_len(self.n_pairs, self.rank, self.world)
        assert n == mine
        if mine:
            self._send[:mine * RECORD_BYTES].copy_(device_bytes_tensor(ptr, mine * RECORD_BYTES, self.device),
                                                   non_blocking=True)
        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        return self._recv

    # ---- gather fused into the kernel: no collective on the data path ----------------------------------------
    def attach_fused_gather(self, n_global_pairs: int, root_only: bool = False) -> bool:
        """Exchange CUDA IPC handles of per-rank whole-batch buffers once; afterwards every ``run`` writes its
        records straight into all ranks' buffers from the kernel epilogue (peer stores over NVLink) — or, with
        ``root_only``, into rank 0's buffer alone (what ONE host-side pose-graph update needs; the other ranks then
        allocate nothing).  Returns False — on EVERY rank, after an agreement round — when some rank could not map a
        peer's buffer (no peer access between the devices); the caller then gathers with :meth:`gather_device` (NCCL)."""
        import torch
        import torch.distributed as dist
        ok = 1
        self._root_only = bool(root_only)
        try:
            self.sm.gather_set_root_only(root_only)
            mine = (self.sm.gather_declare(n_global_pairs) if (root_only and self.rank != 0)
                    else self.sm.gather_export(n_global_pairs))
        except Exception:
            mine, ok = b"\0" * 64, 0
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=self.group)
        if ok:
            try:
                self.sm.gather_attach(handles, self.rank)
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            try:
                self.sm.gather_detach()
            except Exception:
                pass
            self._fused = 0
            return False
        self._fused = int(n_global_pairs)
        dist.barrier(group=self.group)
        return True

    def detach_fused_gather(self):
        import torch.distributed as dist
        self.sm.synchronize()
        dist.barrier(group=self.group)        # nobody may still be storing into a buffer that is about to close
        self.sm.gather_detach()
        self._fused = 0

    def fused_records(self, n_pairs: Optional[int] = None) -> Optional[np.ndarray]:
        """All records in global pair order from this rank's own buffer (rank 0's only after a root-only attach: the
        other ranks get None).  Synchronises this rank's stream and then all ranks (a peer's stores are complete once
        its kernel has finished); a second barrier after the copy keeps a faster rank's NEXT run from storing new
        records into a buffer that is still being read."""
        import torch.distributed as dist
        n = self.n_pairs if n_pairs is None else int(n_pairs)
        self.sm.synchronize()
        dist.barrier(group=self.group)
        out = None
        if not getattr(self, "_root_only", False) or self.rank == 0:
            out = self.sm.gather_fetch(n)
        dist.barrier(group=self.group)
        return out

    def gather(self) -> np.ndarray:
        """All ``n_pairs`` records in global pair order, on the host, on every rank."""
        host = self.gather_device().cpu().numpy()
        return interleave(host.view(RESULT_DTYPE), self.n_pairs, self.world)
