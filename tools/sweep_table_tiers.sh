#!/bin/bash
# Sweep of the table plan's first-tier sizes on the bench workload (run on a GPU box): prints plan-kernel ms / step ms per setting.
out=gpurun_out/sweep_table_tiers.txt
: > $out
for m0 in 384 512 768 1024 2048; do
  for lv0 in 128 192 320; do
    SCONE_TABLE_M0=$m0 SCONE_TABLE_LV0=$lv0 timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 --warmup 3 --e2e-steps 2 > /tmp/sw.json 2>/tmp/sw.err
    python - "$m0" "$lv0" >> $out <<'PY'
import json, sys
try:
    d = json.loads(open('/tmp/sw.json').read().strip().splitlines()[-1])
    k = d['roofline']['kernels']
    print('M0', sys.argv[1], 'LV0', sys.argv[2], 'plan_ms', round(k['cone']['avg_ms'], 3), 'traj_ms', round(k['layer_bwd']['avg_ms'], 3), 'step_ms', round(d['ms_per_step'], 3),
          'retries', d['fused_info']['retries_last_chunk'])
except Exception as e:
    print('M0', sys.argv[1], 'LV0', sys.argv[2], 'failed', e)
PY
  done
done
cat $out
