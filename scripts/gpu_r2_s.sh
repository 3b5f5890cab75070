#!/bin/bash
# second pass of the dataflow wave / lane / grid sweep with the new default (half of the CTA slots per kernel)
O=gpurun_out/r2s; mkdir -p $O; rm -f $O/*
run() { name=$1; shift
  for w in cfg5 cfg3; do
    env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${w}_$name.json 2> $O/${w}_$name.err
  done
}
run new_default TB_NOOP=1
run new_default_again TB_NOOP=1
run w512_l2_g296 TB_WAVE=512 TB_LANES=2 TB_DF_GRID=296
run w512_l2_g148 TB_WAVE=512 TB_LANES=2 TB_DF_GRID=148
run w342_l3_g148 TB_WAVE=342 TB_LANES=3 TB_DF_GRID=148
run w128_l4_g148 TB_WAVE=128 TB_LANES=4 TB_DF_GRID=148
run w256_l4_g222 TB_DF_GRID=222
timeout 900 python -m pytest tests/test_dataflow.py tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -2
tail -c 200 $O/*.err | tail -6
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'launches', d['launches_per_step'], 'frac', round(d['roofline'].get('frac') or 0,3), d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
