#!/bin/bash
# development aid: default bench without the CPU legs; prints the headline, e2e and the hinted re-alignment figure
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-latency > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo rc=$?; tail -3 gpurun_out/bench_h.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_h.json") if l.startswith("{")][0])
print(d["value"], d["e2e"]["value"], d["e2e"]["steps"])
print(d["also"]["corridor_realigned_with_cost_hints"])
PY
