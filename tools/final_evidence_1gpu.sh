#!/bin/bash
# Round-end evidence on ONE B200 (one gpurun call): tests, the contract bench line, launch list, ncu captures, side benches.
TAG=${1:-r2z}
mkdir -p gpurun_out
echo "== pytest"; python -m pytest tests -m gpu -q 2>&1 | tail -2
echo "== bench wide"; python bench.py > gpurun_out/${TAG}_bench_wide.json 2> gpurun_out/${TAG}_bench_wide.err; tail -c 300 gpurun_out/${TAG}_bench_wide.json
echo "== launch list"; ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_wide.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-map-stage --no-parity-check > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; tail -c 200 gpurun_out/${TAG}_bench_under_ncu.log
echo "== ncu wide"; ncu --set full --clock-control none --import-source on -k "regex:k_leaf_hash|k_pass" -c 5 -o gpurun_out/${TAG}_wide \
   python tools/prof_commit.py --ncols 256 --n-log 20 --hash 0 --iters 1 2>&1 | tail -1
echo "== ncu config1"; ncu --set full --clock-control none --import-source on -c 8 -o gpurun_out/${TAG}_config1 \
   python tools/prof_commit.py --ncols 135 --n-log 14 --hash 0 --iters 1 2>&1 | tail -1
echo "== bench config1"; python bench.py --n-log 14 --ncols 135 --steps 20 --no-map-stage > gpurun_out/${TAG}_bench_config1.json 2>/dev/null; tail -c 200 gpurun_out/${TAG}_bench_config1.json
echo "== bench poseidon2"; python bench.py --hash poseidon2 --no-map-stage --no-cpu-baseline > gpurun_out/${TAG}_bench_wide_poseidon2.json 2>/dev/null; tail -c 200 gpurun_out/${TAG}_bench_wide_poseidon2.json
echo "== reference arm (2^18-row sample)"; python bench.py --impl reference --cpu-sample-log 18 --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cat gpurun_out/${TAG}_bench_reference.json | cut -c1-400
echo "== fri"; python tools/fri_bench.py 2>&1 | tail -3; python tools/fri_breakdown.py | tail -1
echo "== quick"; python tools/quick_bench.py 2>&1 | tail -9
