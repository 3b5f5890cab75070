#!/bin/bash
# per-node hybrid roofline of every workload
O=gpurun_out/r2x; mkdir -p $O; rm -f $O/*
for w in cfg2 cfg5 cfg3; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/$w.json 2> $O/$w.err
done
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-other-configs --no-e2e --cpu-budget 0 > $O/cfg4.json 2> $O/cfg4.err
tail -c 300 $O/*.err | tail -5
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2x/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        h=d['roofline']['hybrid']
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'dpx whole', round(d['roofline']['frac_whole_step'],3), 'hybrid', {k:(round(v,4) if isinstance(v,float) else v) for k,v in h.items() if k!='note'})
    except Exception as e: print(f,'ERR',e)
PY
