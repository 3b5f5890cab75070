#!/bin/bash
# executors / kernel instances against the round-1 build: default (per-call choice), forced dataflow, forced level-sync
mkdir -p gpurun_out/r2e; rm -f gpurun_out/r2e/*
B="--no-cpu-baseline --no-e2e --no-other-configs"
run() { # name workload extra-args
  timeout 300 python bench.py --workload $2 $3 --steps 20 --warmup 5 $B > gpurun_out/r2e/auto_$1.json 2> gpurun_out/r2e/auto_$1.err
  TB_DATAFLOW=1 timeout 300 python bench.py --workload $2 $3 --steps 20 --warmup 5 $B > gpurun_out/r2e/df_$1.json 2> gpurun_out/r2e/df_$1.err
  TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload $2 $3 --steps 20 --warmup 5 $B > gpurun_out/r2e/ls_$1.json 2> gpurun_out/r2e/ls_$1.err
  (cd ab_r1 && timeout 300 python bench.py --workload $2 $3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ../gpurun_out/r2e/r1_$1.json 2> ../gpurun_out/r2e/r1_$1.err)
}
run cfg2 cfg2 ""
run cfg3 cfg3 ""
run cfg5 cfg5 ""
timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 $B > gpurun_out/r2e/auto_cfg4.json 2> gpurun_out/r2e/auto_cfg4.err
(cd ab_r1 && timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ../gpurun_out/r2e/r1_cfg4.json 2> ../gpurun_out/r2e/r1_cfg4.err)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2e/pytest.txt; cat gpurun_out/r2e/pytest.txt
tail -c 300 gpurun_out/r2e/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],3),'median',round(d['ms_per_step_median_rank0'],3),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',round(r.get('frac') or 0,3),'1lane',r.get('frac_single_lane'),'share',{k:round(v,2) for k,v in r['share_of_step'].items()})
    except Exception as e: print(f,'ERR',e)
PY
