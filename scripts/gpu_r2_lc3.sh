#!/bin/bash
# with cached work lists the resident step of light calls is device-bound: wave size / lanes / CTAs per dataflow kernel again
O=gpurun_out/r2lc3; mkdir -p $O; rm -f $O/*
run() { # name, env...
  name=$1; shift
  for w in $WLS; do
    env "$@" timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-other-configs --no-e2e --no-cpu-baseline > $O/${w}_$name.json 2> $O/${w}_$name.err
  done
}
WLS="cfg5 cfg3"
run base TB_NOOP=1
run w1024_l1 TB_WAVE=1024 TB_LANES=1
run w1024_l1_g296 TB_WAVE=1024 TB_LANES=1 TB_DF_GRID=296
run w512_l2 TB_WAVE=512 TB_LANES=2
run w512_l2_g296 TB_WAVE=512 TB_LANES=2 TB_DF_GRID=296
run w256_l4_g296 TB_DF_GRID=296
run w128_l8 TB_WAVE=128 TB_LANES=8
WLS="cfg5"
run w342_l3 TB_WAVE=342 TB_LANES=3
run w256_l4_g100 TB_DF_GRID=100
tail -c 200 $O/*.err | tail -8
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2lc3/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'median', round(d.get('ms_per_step_median_rank0',0),4), 'launches', d['launches_per_step'], 'frac', round(d['roofline'].get('frac') or 0,3), d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
