#!/bin/bash
# validation call 3 (1 GPU): full GPU test-suite with the final library, smoke, default bench exit code, cfg4 line
TAG=${1:-v12c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_cfg2_n1.out 2> $OUT/bench_cfg2_n1.err; echo "exit $?" | tee $OUT/bench_cfg2_n1.rc
tail -c 300 $OUT/bench_cfg2_n1.out; tail -3 $OUT/bench_cfg2_n1.err
echo "== bench cfg4"; timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > $OUT/bench_cfg4.out 2> $OUT/bench_cfg4.err; echo "exit $?"
tail -c 300 $OUT/bench_cfg4.out
ls -la $OUT
