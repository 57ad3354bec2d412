#!/bin/bash
# Sweep of the table plan's CTA sizes (first / second tier) on the bench workload (run on a GPU box).
out=gpurun_out/sweep_table_threads.txt
: > $out
for t0 in 64 128; do
  for t1 in 512 1024; do
    SCONE_TABLE_T0=$t0 SCONE_TABLE_T1=$t1 timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 --warmup 3 --e2e-steps 2 > /tmp/sw.json 2>/tmp/sw.err
    python - "$t0" "$t1" >> $out <<'PY'
import json, sys
try:
    d = json.loads(open('/tmp/sw.json').read().strip().splitlines()[-1])
    k = d['roofline']['kernels']
    print('T0', sys.argv[1], 'T1', sys.argv[2], 'plan_ms', round(k['cone']['avg_ms'], 3), 'traj_ms', round(k['layer_bwd']['avg_ms'], 3), 'step_ms', round(d['ms_per_step'], 3))
except Exception as e:
    print('T0', sys.argv[1], 'T1', sys.argv[2], 'failed', e)
PY
  done
done
cat $out
