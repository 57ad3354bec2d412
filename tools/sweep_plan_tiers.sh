# tier-0 table sizes of the fused plan kernel: HS0 LV0 EC0 -> ms per step, per-kernel ms, retries handed to tier 1
for cfg in "1024 192 3072" "2048 192 3072" "1024 128 2048" "4096 256 4096"; do
  set -- $cfg
  SCONE_FUSED_HS0=$1 SCONE_FUSED_LV0=$2 SCONE_FUSED_EC0=$3 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/sw.json"))
k=d["roofline"]["kernels"]; i=d["fused_info"]
print("$cfg", round(d["ms_per_step"],3), "cone", round(k["cone"]["avg_ms"],3), "traj", round(k["layer_bwd"]["avg_ms"],3), "smem0", i["plan_smem_tier0_kb"], "retries", i["retries_last_chunk"])
EOF
done
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu.log 2>&1; python tools/launch_shares.py gpurun_out/r2n_launches.csv
