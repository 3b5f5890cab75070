#!/bin/bash
# wave / lane sweep on the small-branch workload (cfg5) and on cfg2
TAG=${1:-s05}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if ! timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_plan or golden or int16" 2>&1 | tail -3 | grep -q " passed"; then echo "SANITY FAILED"; exit 1; fi
for WL in cfg5 cfg2; do
for combo in "128 4 2" "256 4 1" "512 2 1" "1024 1 1" "256 2 2" "512 4 1" "64 4 4"; do
  set -- $combo
  TB_WAVE=$1 TB_LANES=$2 TB_WAVES_PER_LANE=$3 timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'workload': '$WL', 'wave': $1, 'lanes': $2, 'waves_per_lane': $3, 'ms_per_step': d['ms_per_step'], 'launches_per_step': d['launches_per_step'], 'value': d['value']}))" | tee -a $OUT/sweep.jsonl
done
done
