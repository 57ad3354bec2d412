# tier-0 table sizes of the fused plan kernel: HS0 LV0 EC0 LC0 -> ms per step, per-kernel ms, retries handed to tier 1
for cfg in "1024 192 3072 256" "1024 192 3072 512" "2048 192 3072 512" "1024 128 2048 256"; do
  set -- $cfg
  SCONE_FUSED_HS0=$1 SCONE_FUSED_LV0=$2 SCONE_FUSED_EC0=$3 SCONE_FUSED_LC0=$4 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/sw.json"))
k=d["roofline"]["kernels"]; i=d["fused_info"]
print("$cfg", round(d["ms_per_step"],3), "cone", round(k["cone"]["avg_ms"],3), "traj", round(k["layer_bwd"]["avg_ms"],3), "smem0", i["plan_smem_tier0_kb"], "retries", i["retries_last_chunk"])
EOF
done
