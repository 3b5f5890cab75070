#!/bin/bash
# Short GPU iteration: sanity gate, full GPU parity suite, bench (resident + e2e), 1-lane shares, kernel-phase accounting.
TAG=${1:-q01}; WL=${2:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if ! timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_plan or golden or int16" 2>&1 | tail -5 | tee $OUT/sanity.log | grep -q " passed"; then
  echo "SANITY FAILED - aborting"; cat $OUT/sanity.log; exit 1
fi
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== bench $WL"; timeout 900 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_$WL.json
echo "== 1 lane"; TB_LANES=1 timeout 600 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_${WL}_1lane.json
if [ -f tensorbranching.jl_b200/libtbcuda_kprof.so ]; then
  echo "== kprof"; TBCUDA_LIB=$PWD/tensorbranching.jl_b200/libtbcuda_kprof.so timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "KPROF" | tee $OUT/kprof.log
fi
