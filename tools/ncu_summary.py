"""Development aid: profiles/<name>.json + <name>_details.csv from an `ncu --set full` report of the stage launches of one
bench step (tools/gpu_ncu_final.sh).  Run in the build container: `python tools/ncu_summary.py REPORT OUT_PREFIX "note"`."""
import csv
import json
import subprocess
import sys

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_units=True):
    v = r[col[name]].replace(",", "")
    if v in ("", "n/a"):
        return None
    x = float(v)
    u = units[col[name]]
    if scale_units:
        x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return x


stages = []
for r in body:
    st = {
        "kernel": r[col["Kernel Name"]].replace("void ", "").split("(")[0],
        "grid": int(val(r, "launch__grid_size")), "block": int(val(r, "launch__block_size")),
        "cluster_size": int(val(r, "launch__cluster_size")),
        "registers_per_thread": int(val(r, "launch__registers_per_thread")),
        "duration_ms": val(r, "gpu__time_duration.sum"),
        "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
        "issue_active_per_cycle_active": val(r, "smsp__issue_active.avg.per_cycle_active"),
        "smsp_cycles_active_avg": val(r, "smsp__cycles_active.avg"), "sm_cycles_elapsed_max": val(r, "sm__cycles_elapsed.max"),
        "pipe_alu_pct_active": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_cycles_pct_active": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "pipe_fp64_pct_active": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "eligible_warps_per_cycle": val(r, "smsp__warps_eligible.avg.per_cycle_active"),
        "active_warps_per_scheduler": val(r, "smsp__warps_active.avg.per_cycle_active"),
        "instructions_executed": int(val(r, "smsp__inst_executed.sum")),
    }
    for k in ("barrier", "wait", "short_scoreboard", "math_pipe_throttle", "not_selected"):
        st["stall_" + k] = val(r, f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio")
    stages.append(st)

summary = {
    "source": note,
    "stages": stages,
    "step_duration_ms_under_ncu": sum(s["duration_ms"] for s in stages),
    "dram_bytes_per_launch": sum(s["dram_bytes_read"] + s["dram_bytes_write"] for s in stages),
    "dram_note": "per step (all stage launches of one step; cold-cache, serialised ncu replays)",
    "instructions_executed": sum(s["instructions_executed"] for s in stages),
    "instructions_note": "smsp__inst_executed.sum, one step = the consecutive icp_pairs_kernel launches of the chain",
}
json.dump(summary, open(out + ".json", "w"), indent=1)
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True, check=True).stdout
open(out + "_details.csv", "w").write(det)
print(json.dumps({k: v for k, v in summary.items() if k != "stages"}, indent=1))
for s in stages:
    print(s["kernel"], s["grid"], s["block"], f'{s["duration_ms"]:.3f} ms', f'{s["instructions_executed"] / 1e9:.2f} G inst',
          f'issue {s["issue_active_per_cycle_active"]}', f'alu {s["pipe_alu_pct_active"]:.1f}%')
