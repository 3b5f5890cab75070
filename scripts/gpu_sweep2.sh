#!/bin/bash
TAG=${1:-s02}; WL=${2:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if ! timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_plan or golden or int16" 2>&1 | tail -3 | grep -q " passed"; then echo "SANITY FAILED"; exit 1; fi
for vt in i16 i32; do
for tgt in 0 2 3 4 5; do
  echo "== $vt split target $tgt"
  TB_SPLIT_TARGET=$tgt timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --value-type $vt 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'vt': '$vt', 'split_target': $tgt, 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'launches': d['launches_per_step'], 'share': d['roofline']['share_of_step'], 'frac': d['roofline']['frac']}))" | tee -a $OUT/sweep.jsonl
done
done
