#!/bin/bash
# fused-subtree kernel: shared-memory carveout (residency) A/B
O=gpurun_out/r2u; mkdir -p $O; rm -f $O/*
run() { name=$1; shift
  for w in cfg5 cfg2 cfg3; do
    env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${w}_$name.json 2> $O/${w}_$name.err
  done
}
run base TB_NOOP=1
run c100 TB_FUSED_CARVEOUT=100
run c50 TB_FUSED_CARVEOUT=50
run base2 TB_NOOP=1
run c100b TB_FUSED_CARVEOUT=100
grep -h "fused-subtree kernel" $O/*.err | sort | uniq -c
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2u/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline'].get('frac') or 0,3), d.get('agrees_with_golden'), 'fused share', d['roofline'].get('share_of_step',{}).get('fused'))
    except Exception as e: print(f,'ERR',e)
PY
