#!/bin/bash
# Round-12 gpurun call: parity (incl. index slicing), cfg4 / cfg3 / cfg5 / cfg2 bench lines, ncu evidence for cfg4.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round12.sh [tag]
TAG=${1:-r12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/host_cores.txt; free -g >> $OUT/host_cores.txt
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== sanity"
if ! timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_index_slicing.py -m gpu -x -q -k "single_plan or golden or int16 or sliced" 2>&1 | tail -8 | tee $OUT/sanity.log | grep -q " passed"; then
  echo "SANITY FAILED - aborting"; cat $OUT/sanity.log; exit 1
fi
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench cfg4"; timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_cfg4.json
echo "== bench cfg3 (index sliced 2^3)"; timeout 600 python bench.py --workload cfg3 --steps 20 --warmup 5 2>&1 | tail -2 | tee $OUT/bench_cfg3.json
echo "== bench cfg3 unsliced"; timeout 600 python bench.py --workload cfg3 --slice-k 0 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg3_k0.json
echo "== bench cfg5"; timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 5 2>&1 | tail -2 | tee $OUT/bench_cfg5.json
echo "== bench cfg2"; timeout 900 python bench.py 2>&1 | tail -2 | tee $OUT/bench_cfg2.json
echo "== bench reference cfg4"; timeout 600 python bench.py --impl reference --workload cfg4 --steps 1 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_cfg4.json
echo "== ncu launches cfg4 (8 branches)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_cfg4.csv \
    python bench.py --workload cfg4 --max-branches 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
echo "== ncu full (gemm, cfg4)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 12 -c 6 -o $OUT/prof_gemm_cfg4 \
    python bench.py --workload cfg4 --max-branches 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
ls -la $OUT
