#!/bin/bash
# validation after the one-compile index slicing (1 GPU): GPU test-suite, smoke, cfg3 sliced (e2e through tb_contract_sliced), default bench
TAG=${1:-v12e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -14 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench cfg3"; timeout 600 python bench.py --workload cfg3 > $OUT/bench_cfg3.out 2> $OUT/bench_cfg3.err; echo "exit $?"
echo "== bench cfg3 k=6"; timeout 600 python bench.py --workload cfg3 --slice-k 6 --no-cpu-baseline > $OUT/bench_cfg3_k6.out 2> $OUT/bench_cfg3_k6.err; echo "exit $?"
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_cfg2.out 2> $OUT/bench_cfg2.err; echo "exit $?" | tee $OUT/bench_cfg2.rc
tail -c 300 $OUT/bench_cfg2.out
ls -la $OUT
