#!/bin/bash
# One gpurun call: peak microbench, GPU parity tests, smoke, bench, ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag] [workload]
TAG=${1:-r01}
WL=${2:-cfg2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/host_cores.txt; lscpu | grep -E "Model name|Flags" | cut -c1-400 >> $OUT/host_cores.txt
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== dpx_peak"; timeout 120 ./tensorbranching.jl_b200/dpx_peak | tee $OUT/dpx_peak.json
echo "== sanity (short timeout: a hang here aborts the round instead of burning GPU minutes)"
if ! timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_plan or golden or int16" 2>&1 | tail -5 | tee $OUT/sanity.log | grep -q " passed"; then
  echo "SANITY FAILED - aborting"; cat $OUT/sanity.log; exit 1
fi
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench cfg1"; timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_cfg1.json
echo "== bench $WL"; timeout 1500 python bench.py --workload $WL --steps 5 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_$WL.json
echo "== bench $WL int32 value type"; timeout 900 python bench.py --workload $WL --steps 5 --warmup 3 --value-type i32 2>&1 | tail -1 | tee $OUT/bench_${WL}_i32.json
echo "== bench $WL direct epilogue"; TB_EPI_DIRECT=1 timeout 600 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --value-type i32 2>&1 | tail -1 | tee $OUT/bench_${WL}_epidirect.json
echo "== bench $WL gemm v1"; TB_GEMM_V1=1 timeout 900 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --value-type i32 2>&1 | tail -1 | tee $OUT/bench_${WL}_gemmv1.json
echo "== bench $WL 1 lane"; TB_LANES=1 timeout 900 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_${WL}_1lane.json
echo "== bench reference"; timeout 900 python bench.py --impl reference --workload $WL --steps 2 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_ref_$WL.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
echo "== ncu full (gemm)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 24 -c 4 -o $OUT/prof_gemm \
    python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
ls -la $OUT
