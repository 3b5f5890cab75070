#!/bin/bash
# other_configs inside the default command vs the same workloads run alone (is anything carried over from cfg4?)
O=gpurun_out/r2j; mkdir -p $O; rm -f $O/*
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-budget 3 > $O/default_short.json 2> $O/default_short.err
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-other-configs --cpu-budget 3 > $O/alone_cfg2.json 2> $O/alone_cfg2.err
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-other-configs --cpu-budget 3 > $O/alone_cfg5.json 2> $O/alone_cfg5.err
tail -c 300 $O/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1],'ms',round(d['ms_per_step'],4),'steps',d.get('steps_timed'),'e2e',d['e2e']['ms_per_step'],d['e2e']['steps'],d['e2e']['host_breakdown_rank0'])
        for k,v in d.get('other_configs',{}).items(): print('   ',k,'ms',round(v['ms_per_step'],4),'steps',v['steps'],'frac',v['kernel_frac_of_dpx_peak'],'e2e',v['e2e'])
    except Exception as e: print(f,'ERR',e)
PY
