#!/bin/bash
# 2-GPU validation bundle (one gpurun --gpus 2 call)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== pytest sharded"; python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -8
echo "== pcie"; $TR tools/pcie_bw.py 2>&1 | grep -v "^W\|^\*" | tail -8
echo "== stall hunt peer (resident)"; $TR tools/peer_stall.py peer 60 2>&1 | grep "rank" | tail -30
echo "== stall hunt peer (with pinned copies)"; $TR tools/peer_stall.py peer 30 20 256 1 2>&1 | grep "rank" | tail -30
echo "== stall hunt nccl"; $TR tools/peer_stall.py nccl 30 2>&1 | grep "rank" | tail -12
echo "== bench nccl"; $TR bench.py --gpus 2 --steps 3 --warmup 3 --exchange nccl > gpurun_out/r2c_bench_2gpu_nccl.json 2> gpurun_out/r2c_bench_2gpu_nccl.err; tail -c 1500 gpurun_out/r2c_bench_2gpu_nccl.json; tail -3 gpurun_out/r2c_bench_2gpu_nccl.err
echo "== bench peer"; $TR bench.py --gpus 2 --steps 3 --warmup 3 --exchange peer > gpurun_out/r2c_bench_2gpu_peer.json 2> gpurun_out/r2c_bench_2gpu_peer.err; tail -c 1500 gpurun_out/r2c_bench_2gpu_peer.json; tail -3 gpurun_out/r2c_bench_2gpu_peer.err
