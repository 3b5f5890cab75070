#!/bin/bash
# wave-size / lane-count sweep of the resident-plan bench (one JSON line per combination)
TAG=${1:-s01}; WL=${2:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if ! timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_plan or golden" 2>&1 | tail -3 | grep -q " passed"; then echo "SANITY FAILED"; exit 1; fi
for combo in "64 4" "64 8" "32 8" "128 4" "128 8" "256 4" "16 8"; do
  set -- $combo
  echo "== wave $1 lanes $2"
  TB_WAVE=$1 TB_LANES=$2 timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'wave': $1, 'lanes': $2, 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'launches': d['launches_per_step']}))" | tee -a $OUT/sweep.jsonl
done
