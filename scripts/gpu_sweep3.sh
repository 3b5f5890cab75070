#!/bin/bash
TAG=${1:-s04}; WL=${2:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if ! timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_plan or golden or int16" 2>&1 | tail -3 | grep -q " passed"; then echo "SANITY FAILED"; exit 1; fi
for combo in "2 4 64" "1 4 64" "1 8 64" "1 8 32" "2 8 32" "1 6 48"; do
  set -- $combo
  echo "== gemm ctas/SM $1 lanes $2 wave $3"
  TB_GEMM_CTAS=$1 TB_LANES=$2 TB_WAVE=$3 timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'gemm_ctas': $1, 'lanes': $2, 'wave': $3, 'ms_per_step': d['ms_per_step'], 'value': d['value']}))" | tee -a $OUT/sweep.jsonl
done
