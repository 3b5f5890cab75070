#!/bin/bash
# short reductions as generic steps instead of GEMM tiles: TB_GEMM_MIN_NK sweep
O=gpurun_out/r2q; mkdir -p $O; rm -f $O/*
for nk in 1 2 3 4 5; do
  export TB_GEMM_MIN_NK=$nk
  for w in cfg5 cfg2 cfg3; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${w}_nk$nk.json 2> $O/${w}_nk$nk.err
  done
done
tail -c 300 $O/*.err | tail -12
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'launches', d['launches_per_step'], 'frac', round(d['roofline'].get('frac') or 0,3), d.get('agrees_with_golden'), d['roofline'].get('share_of_step'))
    except Exception as e: print(f,'ERR',e)
PY
