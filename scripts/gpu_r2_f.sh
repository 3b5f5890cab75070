#!/bin/bash
# dataflow kernels sharing the SMs across lanes (grid = slots / lanes): cfg5 / cfg3 / cfg2, forced dataflow vs level-sync
mkdir -p gpurun_out/r2f; rm -f gpurun_out/r2f/*
B="--no-cpu-baseline --no-e2e --no-other-configs"
for wl in cfg5 cfg2 cfg3; do
  for g in 0 148 74 37; do
    TB_DF_GRID=$g TB_DATAFLOW=1 timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 $B > gpurun_out/r2f/df_${wl}_g$g.json 2> gpurun_out/r2f/df_${wl}_g$g.err
  done
  TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 $B > gpurun_out/r2f/ls_$wl.json 2> gpurun_out/r2f/ls_$wl.err
done
timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 $B > gpurun_out/r2f/auto_cfg4.json 2> gpurun_out/r2f/auto_cfg4.err
tail -c 300 gpurun_out/r2f/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],3),'median',round(d['ms_per_step_median_rank0'],3),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',round(r.get('frac') or 0,3),'share',{k:round(v,2) for k,v in r['share_of_step'].items()})
    except Exception as e: print(f,'ERR',e)
PY
