#!/bin/bash
# N-way comparison with extra bench args: gpu_abn2.sh tag "bench args" "ENV1" "ENV2" ...
TAG=$1; ARGS="$2"; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for E in "$@"; do
  env $E timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $ARGS 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'env': '$E', 'ms_per_step': round(d['ms_per_step'], 3), 'value': round(d['value'], 1), 'share': d['roofline']['share_of_step']}))" | tee -a $OUT/ab.jsonl
done
