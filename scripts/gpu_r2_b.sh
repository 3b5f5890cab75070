#!/bin/bash
# round 2: GPU tests, the driver's default bench command (short), and an A/B of the executors against the round-1 build
mkdir -p gpurun_out/r2b
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b/pytest.txt; cat gpurun_out/r2b/pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-budget 5 > gpurun_out/r2b/bench_default.json 2> gpurun_out/r2b/bench_default.err
for rep in 1 2; do
  if [ -d ab_r1 ]; then (cd ab_r1 && timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ../gpurun_out/r2b/r1_$rep.json 2> ../gpurun_out/r2b/r1_$rep.err); fi
  TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2b/ls_$rep.json 2> gpurun_out/r2b/ls_$rep.err
  timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2b/df_$rep.json 2> gpurun_out/r2b/df_$rep.err
done
tail -c 400 gpurun_out/r2b/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],3),'median',round(d['ms_per_step_median_rank0'],3),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',r.get('frac'),'1lane',r.get('frac_single_lane'),'share',{k:round(v,2) for k,v in r['share_of_step'].items()}, 'e2e', d.get('e2e',{}).get('value'), d.get('agrees_with_golden'), d.get('cpu_baseline',{}).get('agrees_with_gpu'))
        for k,v in d.get('other_configs',{}).items(): print('   ',k,{q:(round(x,3) if isinstance(x,float) else x) for q,x in v.items() if q!='workload'})
    except Exception as e: print(f,'ERR',e)
PY
