#!/bin/bash
# bisect of the remaining kernel regression against the round-1 build: arena allocator / signal warp (level-sync, cfg2)
mkdir -p gpurun_out/r2d
for rep in 1 2; do
for v in v5 v5a v5b v5c; do
  TBCUDA_LIB=$PWD/ab_$v/libtbcuda.so TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2d/${v}_$rep.json 2> gpurun_out/r2d/${v}_$rep.err
done
(cd ab_r1 && timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ../gpurun_out/r2d/r1_$rep.json 2> ../gpurun_out/r2d/r1_$rep.err)
done
tail -c 300 gpurun_out/r2d/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],3),'median',round(d['ms_per_step_median_rank0'],3),'share',{k:round(v,2) for k,v in r['share_of_step'].items()})
    except Exception as e: print(f,'ERR',e)
PY
