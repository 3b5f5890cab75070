#!/bin/bash
# final validation (1 GPU): GPU test-suite, smoke, default bench (exit code), cfg5 with weight-adaptive waves, reference arm
TAG=${1:-v12d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_cfg2.out 2> $OUT/bench_cfg2.err; echo "exit $?" | tee $OUT/bench_cfg2.rc
tail -c 300 $OUT/bench_cfg2.out
echo "== bench cfg5"; timeout 600 python bench.py --workload cfg5 > $OUT/bench_cfg5.out 2> $OUT/bench_cfg5.err; echo "exit $?"
echo "== bench cfg1"; timeout 600 python bench.py --workload cfg1 > $OUT/bench_cfg1.out 2> $OUT/bench_cfg1.err; echo "exit $?"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.out 2> $OUT/bench_ref.err; echo "exit $?"
tail -c 300 $OUT/bench_ref.out
ls -la $OUT
