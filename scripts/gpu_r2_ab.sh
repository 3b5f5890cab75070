#!/bin/bash
# A/B on one box: round-1 build (ab_r1/) vs this tree, level-synchronous and dataflow executors, alternating
mkdir -p gpurun_out/r2ab
for rep in 1 2; do
  (cd ab_r1 && timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ../gpurun_out/r2ab/r1_$rep.json 2> ../gpurun_out/r2ab/r1_$rep.err)
  TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2ab/ls_$rep.json 2> gpurun_out/r2ab/ls_$rep.err
  timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2ab/df_$rep.json 2> gpurun_out/r2ab/df_$rep.err
done
tail -c 300 gpurun_out/r2ab/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2ab/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],3),'median',round(d['ms_per_step_median_rank0'],3),'share',{k:round(v,2) for k,v in d['roofline']['share_of_step'].items()})
    except Exception as e: print(f,'ERR',e)
PY
