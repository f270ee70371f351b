"""CPU: bench.py's reference arm prints ONE JSON line with the contract's keys (the product arm needs a GPU and is
exercised on the B200 box); __graft_entry__.build() is idempotent and leaves every library in place."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, DPGICP_BENCH_CPU_TARGET_S="0.5")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "icp_cov_scan_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_build_entry_point():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    for rel in ("dpg_slam_b200/libdpgicp.so", "dpg_slam_b200/libdpgsynth.so", "dpg_slam_b200/dpg_batch_runner",
                "oracle/libdpgoracle.so"):
        assert os.path.exists(os.path.join(ROOT, rel)), rel
    if os.path.isdir("/root/reference/src/icp_cov"):
        assert os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdpgref.so"))


def test_roofline_helpers_definitions():
    """The two ceilings bench.py quotes: the FP32 peak is the LARGER of the two measured separately-rounded rates (scalar and
    packed chains run at the same rate; the packed probe once reported twice that, from fused instructions) and the
    shared-memory bound charges two write-back cycles per warp-level point evaluation."""
    sys.path.insert(0, ROOT)
    import bench
    peak, scalar, packed = bench.fp32_peak({"mul_add_ops_per_s": 36.2e12, "mul_add_packed_ops_per_s": 37.1e12, "fma_ops_per_s": 36e12}, 37.2)
    assert (peak, scalar, packed) == (37.1, 36.2, 37.1)
    assert bench.fp32_peak({"mul_add_ops_per_s": 0.0}, 37.2)[0] == 37.2            # no probe: nominal rate
    blk = bench.smem_block({"distance_evals": 32 * 1000}, k_ms=1.0, sm_mhz=1000.0, n_sm=2)
    assert blk["busy_cycles_per_launch"] == 2000.0 and blk["available_cycles"] == 2.0e6 and blk["frac"] == 1e-3
