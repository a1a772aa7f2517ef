#!/bin/bash
# One bench.py line at N GPUs with the default flags (what the driver runs): usage tools/scale_bench.sh N [tag]
N=$1; TAG=${2:-r2z}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1]
print("N=$N value %.2f Gelem/s (%.2f ms) cap_xor %s parity %s | e2e %.2f (%.1f ms) all_out %.2f | map %.1f whole %.1f %s" % (
  d["value"], d["ms_per_step"], d["cap_xor"], d["parity_check"].get("cap_equals_single_gpu"), d["e2e"]["value"], d["e2e"]["ms_per_step"],
  d["e2e"]["all_outputs_to_host"]["value"], d["map_stage"]["value"], d["map_stage"]["whole_prover"]["value"],
  [round(x,1) for x in d["map_stage"]["whole_prover"]["runs_proofs_per_s"]]))
PY
tail -2 gpurun_out/${TAG}_bench_${N}gpu.err
