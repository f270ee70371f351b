#!/bin/bash
# development aid: corridor + loop-closure bench for library variants x DPGICP_HANDOVER settings
# usage: VARIANTS="default x" HANDOVERS="_ 1,1" tools/gpu_ab_env.sh
for v in ${VARIANTS:-default}; do if [ $v = default ]; then unset DPGICP_LIBRARY; else export DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_$v.so; fi
for H in ${HANDOVERS:-_}; do if [ "$H" = "_" ]; then unset DPGICP_HANDOVER; else export DPGICP_HANDOVER=$H; fi
echo "== $v handover '$H'"; python bench.py --no-cpu-baseline --no-latency --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); r=d['roofline']; a=d['also']['loop_closure']
print('corridor %.0f pairs/s %.3f ms stages %s | loop %.0f'%(d['value'],d['ms_per_step'],[round(x,2) for x in r['stage_ms']],a['value']))"
done; done
