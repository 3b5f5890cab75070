#!/bin/bash
# e2e pipeline batch-size sweep on cfg2 (host-bound path): TB_BATCH_MAX x host threads
TAG=${1:-s06}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for B in 0 128 256 384; do
  TB_BATCH_MAX=$B timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'batch_max': $B, 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'e2e_value': d['e2e']['value'], 'host': d['e2e']['host_breakdown_rank0']}))" | tee -a $OUT/sweep.jsonl
done
