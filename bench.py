#!/usr/bin/env python
"""bench.py — ICP + covariance scan pairs/s of the scan-matching path (BASELINE.json metric).

A "step" is one pass of the hot path (DpgSLAM::runIcp + calculate_ICP_COV semantics, reference
src/dpg_slam/dpg_slam.cc:362-446) over one batch of synthetic Hokuyo-like scan pairs:

  workload  BASELINE.json configs[1]: 2D corridor trajectory, 5000 sequential odometry scan pairs per
            GPU, 1081 beams/scan, 270 deg FOV, every point enters ICP (downsample divisor 1 = the
            "1081-beam" setting), point-to-point, reciprocal correspondences, covariance = the
            intended Censi form on the final ICP correspondences (CENSI_CORR), all other parameters
            the reference's defaults (500 iterations max, 0.6 m gate, eps 5e-9).
  value     pairs/s with the scan store and pair list already resident in HBM; CUDA events on the
            launching stream, L2 flushed between steps, max over ranks.
  e2e       the same batch through the public API with HOST buffers: raw ranges H2D + on-device
            scan->cloud conversion + pair list H2D + ICP/covariance + records D2H, every step.
  N > 1     pairs sharded round-robin (pair k -> rank k % N), scan store replicated, records
            all-gathered over NCCL inside the step; weak scaling (5000 pairs per GPU).

`--impl reference` times the reference's CPU algorithm on the host cores (the oracle port,
oracle/dpg_oracle.c, OpenMP over pairs): PCL itself cannot be built in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "icp_cov_scan_pairs_per_sec"
UNIT = "pairs/s"
# The headline workload is BASELINE configs[1]; the others are the remaining single-GPU-sized configs, selectable
# with --workload for additional measurements (they are parity-test cases first, not the bench line).
WORKLOADS = {
    "corridor": dict(desc="BASELINE configs[1]: synthetic corridor trajectory, 5000 sequential odometry scan pairs per GPU",
                     pairs_per_gpu=5000, beams=1081, seed=2, metric=0),
    "loop_closure": dict(desc="BASELINE configs[2]: loop-closure candidate sweep, 100k scan pairs over 2000 scans with random "
                              "initial offsets, per GPU", pairs_per_gpu=100_000, beams=1081, seed=3, metric=0),
    "dense": dict(desc="BASELINE configs[3] shape: dense 4096-beam scans, point-to-line, 125k pairs per GPU (1M over 8 GPUs)",
                  pairs_per_gpu=125_000, beams=4096, seed=4, metric=1),
}
WORKLOAD = "corridor"
SEARCH = "pruned"            # --search projective: the approximate beam-order search (north-star extension), not the headline
PAIRS_PER_GPU = WORKLOADS[WORKLOAD]["pairs_per_gpu"]
N_BEAMS = WORKLOADS[WORKLOAD]["beams"]


def select_workload(name: str):
    global WORKLOAD, PAIRS_PER_GPU, N_BEAMS
    WORKLOAD = name
    PAIRS_PER_GPU = WORKLOADS[name]["pairs_per_gpu"]
    N_BEAMS = WORKLOADS[name]["beams"]


# ---- workload ----------------------------------------------------------------------------------------------
def make_workload(n_pairs: int):
    from dpg_slam_b200 import synth
    w = WORKLOADS[WORKLOAD]
    if WORKLOAD == "corridor":
        return synth.config_corridor(n_pairs=n_pairs, n_beams=N_BEAMS, seed=w["seed"])
    if WORKLOAD == "loop_closure":
        return synth.config_loop_closure(n_pairs=n_pairs, n_scans=2000, n_beams=N_BEAMS, seed=w["seed"])
    return synth.config_loop_closure(n_pairs=n_pairs, n_scans=20_000, n_beams=N_BEAMS, seed=w["seed"])


def bench_params():
    from dpg_slam_b200._abi import COV_CENSI_CORR, SEARCH_PROJECTIVE, SEARCH_PRUNED, Params
    return Params.defaults(downsample_divisor=1, cov_mode=COV_CENSI_CORR, metric=WORKLOADS[WORKLOAD]["metric"],
                           search=SEARCH_PROJECTIVE if SEARCH == "projective" else SEARCH_PRUNED)


def workload_config(n_gpus: int, extra=None):
    w = WORKLOADS[WORKLOAD]
    cfg = {"workload": w["desc"],
           "pairs_per_gpu": PAIRS_PER_GPU, "global_pairs": PAIRS_PER_GPU * n_gpus, "beams": N_BEAMS,
           "fov_deg": 270, "downsample_divisor": 1, "metric_kind": "point_to_line" if w["metric"] else "point_to_point",
           "reciprocal": True,
           "cov_mode": "CENSI_CORR(cap 200)", "max_iterations": 500, "max_correspondence_distance_m": 0.6,
           "search": ("exact pruned (bounding-box groups)" if SEARCH == "pruned" else
                      "PROJECTIVE (approximate: beam-order projection, window 8 each side; not the reference's exact search)"),
           "sharding": f"round-robin over {n_gpus} GPU(s), scan store replicated",
           "l2": "flushed between timed steps (512 MiB memset outside the event pair)", "seed": w["seed"]}
    if extra:
        cfg.update(extra)
    return cfg


# ---- clocks --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline / reference arm ---------------------------------------------------------------------------
def cpu_sample_run(wl, pts, off, p, idx, threads=0):
    from oracle import oracle_py as O
    t0 = time.perf_counter()
    rec, used = O.run_batch(pts, off, wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx], p, fast=1, threads=threads)
    return time.perf_counter() - t0, used, rec


def choose_cpu_sample(wl, pts, off, p, target_s=12.0):
    """Bounded sample of the workload: pilot on 4 pairs per core, then size for ~target_s of CPU time."""
    cores = os.cpu_count() or 1
    n = wl.n_pairs
    pilot = np.linspace(0, n - 1, min(n, 4 * cores)).astype(np.int64)
    dt, used, _ = cpu_sample_run(wl, pts, off, p, pilot)
    per_pair = dt / len(pilot)
    m = int(max(len(pilot), min(n, target_s / max(per_pair, 1e-9))))
    return np.linspace(0, n - 1, m).astype(np.int64), used


def oracle_clouds(wl):
    from oracle import oracle_py as O
    return O.clouds_from_ranges(wl.ranges, wl.scanner)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "oracle", "libdpgoracle.so")):
        g.build()
    wl = make_workload(PAIRS_PER_GPU)
    p = bench_params()
    pts, off = oracle_clouds(wl)
    idx, used = choose_cpu_sample(wl, pts, off, p, target_s=float(os.environ.get("DPGICP_BENCH_CPU_TARGET_S", "8.0")))
    for _ in range(max(args.warmup, 0)):
        cpu_sample_run(wl, pts, off, p, idx[:max(8, len(idx) // 8)])
    t_total = 0.0
    for _ in range(args.steps):
        dt, used, _ = cpu_sample_run(wl, pts, off, p, idx)
        t_total += dt
    value = len(idx) * args.steps / t_total
    sample = (f"{len(idx)} of the {wl.n_pairs} pairs per step (evenly strided), oracle port with exact uniform-grid NN, "
              f"OpenMP over pairs on {used} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = CPU restatement of runIcp + calculate_ICP_COV (PCL ICP is un-vendored and cannot be built here)"}
    print(json.dumps(line), flush=True)
    return 0


# ---- product arm ------------------------------------------------------------------------------------------------
def algorithmic_flops(rec, counts_s, counts_t, sum_corr):
    """SURVEY.md §8d normative work: I*(5*Ns*Nt + 8*Ns) + 14*sum K  +  60*N_H + 45*min(K,200)."""
    it = rec["iterations"].astype(np.float64)
    ns, nt = counts_s.astype(np.float64), counts_t.astype(np.float64)
    k_last = rec["n_correspondences"].astype(np.float64)
    icp = float(np.sum(it * (5.0 * ns * nt + 8.0 * ns)) + 14.0 * sum_corr)
    cov = float(np.sum(60.0 * k_last + 45.0 * np.minimum(k_last, 200.0)))
    return icp, cov


def run_product(args):
    import torch
    import torch.distributed as dist
    from dpg_slam_b200 import sharded
    from dpg_slam_b200.scanmatch import ScanMatcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_global = PAIRS_PER_GPU * world
    wl = make_workload(n_global)
    p = bench_params()
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream), ScanMatcher(local) as sm:
        sm.set_stream(stream.cuda_stream)
        shard = sharded.ShardedScanMatcher(sm, rank, world, None, dev)
        idx = sharded.shard_indices(n_global, rank, world)

        # ---------------- resident-input arm (value) ----------------
        sm.upload_ranges(wl.ranges, wl.scanner)
        shard.set_pairs(wl.src_idx, wl.tgt_idx, wl.guess)

        # N > 1: the gather of the records is fused into the kernel epilogue (peer stores into every rank's
        # whole-batch buffer over NVLink); DPGICP_BENCH_GATHER=nccl uses an all-gather collective after the kernel
        fused = world > 1 and os.environ.get("DPGICP_BENCH_GATHER", "fused") != "nccl"
        if fused:
            fused = shard.attach_fused_gather(n_global)       # False on every rank if a peer buffer could not be mapped

        def step_resident():
            shard.run(p)
            if world > 1 and not fused:
                shard.gather_device()

        sampler = ClockSampler(local)          # started before the warm-up so that nvidia-smi is already sampling when
        sampler.start()                        # the timed region begins; stopped after the kernel-only timing loop
        for _ in range(max(args.warmup, 3)):
            step_resident()
        barrier()
        launches0 = sm.last_run_counters()["kernel_launches"]
        evs = []
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(0)                                   # L2 flush, outside the timed event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_resident()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        t_wall = time.perf_counter() - t_wall0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = sm.last_run_counters()["kernel_launches"] - launches0
        counters = sm.last_run_counters()
        rec_local = sm.fetch_results()
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max = float(t.item())
        value = n_global * args.steps / (dev_ms_max * 1e-3)

        # kernel-only duration (no gather) for the roofline of the dominant kernel
        kevs = []
        for _ in range(max(3, min(args.steps, 10))):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); shard.run(p); e1.record(stream)
            kevs.append((e0, e1))
        torch.cuda.synchronize(dev)
        k_ms = float(np.mean([a.elapsed_time(b) for a, b in kevs]))
        clocks = sampler.stop()

        # ---------------- end-to-end arm (host buffers in, host records out) ----------------
        ranges_pin = torch.from_numpy(wl.ranges).pin_memory()
        out_pin = torch.empty(len(idx) * sharded.RECORD_BYTES, dtype=torch.uint8).pin_memory()
        src_l, tgt_l, guess_l = wl.src_idx[idx], wl.tgt_idx[idx], wl.guess[idx]

        used = None
        if world > 1:      # a rank needs only the scans its shard touches: read in place from page-locked memory
            used = shard.plan_subset(wl.src_idx, wl.tgt_idx, wl.n_scans)
            src_e, tgt_e = shard._remap[src_l], shard._remap[tgt_l]
        else:
            src_e, tgt_e = src_l, tgt_l

        def step_e2e():
            if used is None:
                sm.upload_ranges_ptr(ranges_pin.data_ptr(), wl.n_scans, N_BEAMS, wl.scanner) # H2D + scan->cloud
            else:
                sm.upload_ranges_subset(ranges_pin.data_ptr(), used, wl.scanner, n_scans_total=wl.n_scans, n_beams=N_BEAMS)
            sm.set_pairs(src_e, tgt_e, guess_l)                                              # H2D pair list
            sm.run(p)
            if world > 1 and not fused:
                shard.gather_device()
            sm.fetch_results_ptr(out_pin.data_ptr(), len(idx))                               # D2H records (syncs)

        for _ in range(3):
            step_e2e()
        barrier()
        e2e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = n_global * e2e_steps / float(t.item())
        n_up = wl.n_scans if used is None else len(used)
        h2d = int(4 * N_BEAMS * n_up + 24 * len(idx) + (0 if used is None else 4 * n_up))
        d2h = int(sharded.RECORD_BYTES * len(idx) + 4 * n_up + 4)
        e2e_rec = np.frombuffer(out_pin.numpy().tobytes(), dtype=rec_local.dtype)
        assert e2e_rec.tobytes() == rec_local.tobytes(), "e2e records differ from the resident-arm records"

        if fused:
            allrec = shard.fused_records()                                                   # whole batch, global order
            assert allrec[idx].tobytes() == rec_local.tobytes(), "fused gather does not hold this rank's records"
            shard.detach_fused_gather()
        probe = sm.fp32_probe() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel (icp_pairs_kernel), rank 0's launch ----------------
    from dpg_slam_b200._abi import FLAG_CONVERGED
    counts = (wl.ranges < wl.scanner.range_max).sum(axis=1)
    ns, nt = counts[src_l], counts[tgt_l]
    icp_fl, cov_fl = algorithmic_flops(rec_local, ns, nt, counters["correspondences"])
    alg_tflops = (icp_fl + cov_fl) / (k_ms * 1e-3) / 1e12
    exec_tflops = 5.0 * counters["distance_evals"] / (k_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    nominal = 148 * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    measured = probe["mul_add_ops_per_s"] / 1e12
    packed = probe.get("mul_add_packed_ops_per_s", 0.0) / 1e12
    # the distance loop issues its subtractions and multiplications as packed pairs (FADD2 / FMUL2): the FP32 ceiling
    # of separately rounded operations on this GPU is the packed rate
    peak = packed if packed > 0 else (measured if measured > 0 else nominal)
    traffic, ncu_inst = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_final_icp_kernel_ncu.json")))
        if WORKLOAD == "corridor" and SEARCH == "pruned":       # the capture is of this workload's launch
            traffic = prof.get("dram_bytes_per_launch")
            ncu_inst = prof.get("instructions_executed")
    except Exception:
        pass
    alg_bytes = float(np.sum(8.0 * (ns + nt) + 20 + 112))
    roofline = {"kernel": "dpg::icp_pairs_kernel<WARPS,SEARCH,CLUSTER> (persistent CTAs; one step = a chain of up to 4 launches with "
                          "growing warps per pair, the last as 4-CTA clusters, timed together)",
                "bound": "fp32", "achieved": alg_tflops, "peak": peak, "unit": "TFLOP/s", "frac": alg_tflops / peak,
                "peak_source": "measured on this GPU in this run by dpgicp_fp32x2_probe: separately rounded FMUL2+FADD2 chains "
                               "(packed pairs, 2 operations per issue slot; the bit-exact distance loop may not use FMA); "
                               "MEASURED_PEAKS.json has no FP32 figure; "
                               f"nominal scalar rate 148 SM x 128 lanes x {peaks.get('sm_max_mhz', 1965.0)} MHz = {nominal:.1f}",
                "achieved_definition": "ALGORITHMIC brute-force flops (SURVEY 8d: I*(5*Ns*Nt+8*Ns)+14*K + 60*N_H+45*min(K,200)) / kernel time; "
                                       "the exact pruned search skips most of them, so frac can exceed 1 — see executed_*",
                "executed_tflops": exec_tflops, "executed_frac": exec_tflops / peak,
                "executed_definition": "5 flop x distance evaluations actually executed (device counter) / kernel time",
                "kernel_ms": k_ms, "traffic": traffic,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs"), "frac": (alg_bytes / (k_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None},
                "fp32_probe_tops": {"mul_add": measured, "mul_add_packed": packed, "fma": probe["fma_ops_per_s"] / 1e12}}
    if ncu_inst:
        # issue-slot view: warp instructions of one step (ncu capture of this same workload, profiles/) over the
        # live kernel time, against 4 schedulers x 148 SMs x SM clock
        issue_peak = 4.0 * 148 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
        roofline["issue"] = {"warp_instructions_per_step": ncu_inst, "achieved_ginst_s": ncu_inst / (k_ms * 1e-3) / 1e9,
                             "peak_ginst_s": issue_peak / 1e9, "frac": ncu_inst / (k_ms * 1e-3) / issue_peak}

    # ---------------- CPU baseline on this box's host cores (bounded sample; N = 1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        pts, off = oracle_clouds(wl)
        cidx, used = choose_cpu_sample(wl, pts, off, p, target_s=12.0)
        dt, used, cref = cpu_sample_run(wl, pts, off, p, cidx)
        same = all(np.array_equal(cref[f], rec_local[cidx][f]) for f in
                   ("tx", "ty", "iterations", "status", "n_correspondences", "mse"))
        sidx = cidx[::max(1, len(cidx) // 48)][:48]             # the reference itself is single-threaded (ros::spin)
        dt1, _, _ = cpu_sample_run(wl, pts, off, p, sidx, threads=1)
        cpu = {"value": len(cidx) / dt, "unit": UNIT, "cores": used, "kind": "port",
               "single_thread_value": len(sidx) / dt1, "single_thread_sample": f"{len(sidx)} pairs, {dt1:.1f} s",
               "sample": f"{len(cidx)} of {wl.n_pairs} pairs (evenly strided), {dt:.1f} s, oracle port with exact uniform-grid NN, "
                         f"OpenMP over pairs", "records_equal_gpu": bool(same)}

    conv = float(((rec_local["status"] & FLAG_CONVERGED) != 0).mean())
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, {"gather": "none (1 GPU)" if world == 1 else
                                              ("peer stores from the kernel epilogue (fused)" if fused else "NCCL all_gather")}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "timing": "host clock around upload+convert+pairs+ICP/cov+fetch, max over ranks"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall,
            "stats": {"mean_iterations": float(rec_local["iterations"].mean()), "p95_iterations": float(np.percentile(rec_local["iterations"], 95)),
                      "max_iterations": int(rec_local["iterations"].max()),
                      "converged_frac": conv, "distance_evals_per_launch": counters["distance_evals"],
                      "box_tests_per_launch": counters["box_tests"]}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="corridor", choices=sorted(WORKLOADS))
    ap.add_argument("--search", default="pruned", choices=["pruned", "projective"])
    args = ap.parse_args()
    select_workload(args.workload)
    global SEARCH
    SEARCH = args.search
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
