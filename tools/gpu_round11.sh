#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --no-cpu-baseline --workload loop_closure > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo "loop rc=$?"; cut -c1-200 gpurun_out/bench_loop.json
timeout 2400 python bench.py --steps 3 --no-cpu-baseline --workload dense > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; echo "dense rc=$?"; cut -c1-200 gpurun_out/bench_dense.json; tail -3 gpurun_out/bench_dense.err
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from dpg_slam_b200.scanmatch import ScanMatcher
rng = np.random.default_rng(0)
with ScanMatcher(0) as sm:
    for n in (2000, 50000, 200000, 400000):
        side = (n / 4.0) ** 0.5 * 1.0          # ~4 nodes per m^2 -> ~300 neighbours within 5 m
        xy = rng.uniform(0, side * 4, (n, 2)).astype(np.float32)
        ps = (np.arange(n) // (n // 8 + 1)).astype(np.int32)
        t = time.time(); s, tg = sm.enumerate_pairs(xy, ps, 5.0, 2.0); dt = time.time() - t
        print(f"enumerate n={n}: {len(s)} pairs in {dt:.3f} s", flush=True)
PY
