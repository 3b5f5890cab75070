#!/bin/bash
# dataflow executor on light plans (cfg5) and few plans (cfg3): wave size / lanes / CTAs per persistent kernel
O=gpurun_out/r2r; mkdir -p $O; rm -f $O/*
run() { # name, env...
  name=$1; shift
  for w in cfg5 cfg3; do
    env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${w}_$name.json 2> $O/${w}_$name.err
  done
}
run base TB_NOOP=1
run w512_l2 TB_WAVE=512 TB_LANES=2
run w1024_l1 TB_WAVE=1024 TB_LANES=1
run w1024_l4 TB_WAVE=1024 TB_LANES=4 TB_WAVES_PER_LANE=1
run w512_l4 TB_WAVE=512 TB_LANES=4
run w128_l4 TB_WAVE=128 TB_LANES=4
run w256_l4_g148 TB_DF_GRID=148
run w256_l4_g296 TB_DF_GRID=296
run w1024_l1_g296 TB_WAVE=1024 TB_LANES=1 TB_DF_GRID=296
tail -c 200 $O/*.err | tail -8
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'launches', d['launches_per_step'], 'frac', round(d['roofline'].get('frac') or 0,3), d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
