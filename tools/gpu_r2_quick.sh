#!/bin/bash
# quick A/B: GPU suite (fail fast) + the two single-GPU bench workloads without CPU legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_quick.log | cut -c1-200
for V in ${VARIANTS:-default}; do
  LIB=""; [ "$V" != "default" ] && LIB="$PWD/dpg_slam_b200/libdpgicp_$V.so"
  DPGICP_LIBRARY=$LIB timeout 600 python bench.py --no-cpu-baseline --no-latency --steps 20 ${BENCH_ARGS:-} > gpurun_out/bench_quick_$V.json 2> gpurun_out/bench_quick_$V.err; echo "bench $V rc=$?"
  python - "$V" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_quick_{v}.json') if l.startswith('{')][0])
    r=d['roofline']; a=d.get('also',{}).get('loop_closure')
    print(v, 'corridor %.0f pairs/s  %.3f ms  stages %s  e2e %.0f' % (d['value'], d['ms_per_step'], [round(x,2) for x in r['stage_ms']], d['e2e']['value']))
    if a: print(v, 'loop    %.0f pairs/s  %.2f ms  stages %s' % (a['value'], a['ms_per_step'], [round(x,2) for x in a['roofline']['stage_ms']]))
except Exception as e:
    print(v, 'no result', e)
PY
done
