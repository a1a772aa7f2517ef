#!/bin/bash
# N-GPU validation bundle (one gpurun --gpus N call): usage tools/run_2gpu.sh [N] [tag]
N=${1:-2}; TAG=${2:-r2k}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then echo "== pytest sharded"; python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -4; fi
echo "== stall hunt peer (resident)"; $TR tools/peer_stall.py peer 40 2>&1 | grep "rank 0" | tail -8
echo "== stall hunt peer (with pinned copies)"; $TR tools/peer_stall.py peer 20 20 256 1 2>&1 | grep "rank 0" | tail -8
echo "== stall hunt nccl"; $TR tools/peer_stall.py nccl 20 2>&1 | grep "rank 0" | tail -4
for ex in peer nccl; do
  echo "== bench $ex"; $TR bench.py --gpus $N --steps 5 --warmup 3 --exchange $ex > gpurun_out/${TAG}_bench_${N}gpu_$ex.json 2> gpurun_out/${TAG}_bench_${N}gpu_$ex.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_${N}gpu_$ex.json") if l.startswith("{")][-1]
    print("value %.2f ms %.2f cap_xor %s parity %s e2e %.2f (%.1f ms) all_out %.2f map %.1f" % (d["value"], d["ms_per_step"], d["cap_xor"], d["parity_check"].get("cap_equals_single_gpu"), d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["all_outputs_to_host"]["value"], d["map_stage"]["value"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
  tail -3 gpurun_out/${TAG}_bench_${N}gpu_$ex.err
done
