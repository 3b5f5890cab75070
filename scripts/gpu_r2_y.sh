#!/bin/bash
# the default command three times: are the other_configs steady (staging-slot policy)?
O=gpurun_out/r2y; mkdir -p $O; rm -f $O/*
for i in 1 2 3; do
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --cpu-budget 2 > $O/default_$i.json 2> $O/default_$i.err
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'cfg4 ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))
        for k,v in d['other_configs'].items(): print('   ',k,'mean',round(v['ms_per_step'],3),'median',round(v['ms_per_step_median'],3),'max',round(v['ms_per_step_max'],3),'steps',v['steps'],'e2e',round(v['e2e']['ms_per_step'],3), 'hybrid', v.get('whole_step_frac_of_hybrid_roofline'))
    except Exception as e: print(f,'ERR',e)
PY
