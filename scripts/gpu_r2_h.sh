#!/bin/bash
# split-K for parallelism (TB_SPLIT_TARGET = log2 of the tiles a node should have) under the dataflow executor: cfg3 / cfg5
O=gpurun_out/r2h; mkdir -p $O; rm -f $O/*
B="--no-cpu-baseline --no-e2e --no-other-configs"
for wl in cfg3 cfg5; do
  for t in 0 5 6 7 8; do
    TB_SPLIT_TARGET=$t timeout 300 python bench.py --workload $wl --steps 50 --warmup 10 $B > $O/${wl}_t$t.json 2> $O/${wl}_t$t.err
  done
done
TB_SPLIT_TARGET=6 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 $B > $O/cfg2_t6.json 2> $O/cfg2_t6.err
tail -c 300 $O/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d['roofline']
        print(f.split('/')[-1],'ms',round(d['ms_per_step'],4),'median',round(d['ms_per_step_median_rank0'],4),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',round(r.get('frac') or 0,3),d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
