#!/bin/bash
TAG=${1:-v12g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_cfg2.out 2> $OUT/bench_cfg2.err; echo "exit $?" | tee $OUT/bench_cfg2.rc
tail -c 300 $OUT/bench_cfg2.out
ls -la $OUT
