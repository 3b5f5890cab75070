#!/bin/bash
TAG=${1:-c01}; WL=${2:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if ! timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_plan or golden or int16" 2>&1 | tail -3 | grep -q " passed"; then echo "SANITY FAILED"; exit 1; fi
TB_DUMP_LAUNCHES=$OUT/launches_i32.csv timeout 600 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/b_i32.json 2>&1
TB_DUMP_LAUNCHES=$OUT/launches_i16.csv timeout 600 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --value-type i16 > $OUT/b_i16.json 2>&1
tail -c 400 $OUT/b_i16.json
echo "== ncu full k_gemm2h"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm2h -s 24 -c 4 -o $OUT/prof_gemm2h \
    python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --value-type i16 > $OUT/ncu_full.log 2>&1
ls -la $OUT
