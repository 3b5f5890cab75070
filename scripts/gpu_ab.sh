#!/bin/bash
# A/B on one box: alternate two environment settings, N repetitions each.  usage: gpu_ab.sh tag "ENV_A" "ENV_B" [reps] [extra bench args]
TAG=$1; A="$2"; B="$3"; REPS=${4:-3}; EXTRA="$5"
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for r in $(seq 1 $REPS); do
  for v in A B; do
    if [ $v = A ]; then E="$A"; else E="$B"; fi
    env $E timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'variant': '$v', 'env': '$E', 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'share': d['roofline']['share_of_step']}))" | tee -a $OUT/ab.jsonl
  done
done
