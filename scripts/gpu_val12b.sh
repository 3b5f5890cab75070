#!/bin/bash
# validation call 2 (1 GPU): full GPU test-suite, smoke, default bench with exit code and full tail
TAG=${1:-v12b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_cfg2_n1.out 2> $OUT/bench_cfg2_n1.err; echo "exit $?" | tee $OUT/bench_cfg2_n1.rc
tail -c 600 $OUT/bench_cfg2_n1.out; tail -5 $OUT/bench_cfg2_n1.err
echo "== bench default again (exit code under MALLOC_CHECK_)"; MALLOC_CHECK_=3 timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_cfg2_mc.out 2> $OUT/bench_cfg2_mc.err; echo "exit $?" | tee $OUT/bench_cfg2_mc.rc
tail -3 $OUT/bench_cfg2_mc.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.out 2>$OUT/bench_ref.err; echo "exit $?"
echo "== bench cfg5 / cfg3 / cfg1"; for w in cfg5 cfg3 cfg1; do timeout 600 python bench.py --workload $w > $OUT/bench_$w.out 2> $OUT/bench_$w.err; echo "$w exit $?"; done
ls -la $OUT
