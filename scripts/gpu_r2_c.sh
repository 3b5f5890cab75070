#!/bin/bash
# does the lost warp-uniform address arithmetic explain the 12 % regression of the GEMM kernel?  ab_v5 = this tree's
# kernels without the generic-tile call in the consumer warps (level-synchronous executor only)
mkdir -p gpurun_out/r2c
for rep in 1 2; do
  TBCUDA_LIB=$PWD/ab_v5/libtbcuda.so TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2c/v5_cfg2_$rep.json 2> gpurun_out/r2c/v5_cfg2_$rep.err
  TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2c/ls_cfg2_$rep.json 2> gpurun_out/r2c/ls_cfg2_$rep.err
done
TBCUDA_LIB=$PWD/ab_v5/libtbcuda.so TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2c/v5_cfg4.json 2> gpurun_out/r2c/v5_cfg4.err
TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2c/ls_cfg4.json 2> gpurun_out/r2c/ls_cfg4.err
timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2c/df_cfg4.json 2> gpurun_out/r2c/df_cfg4.err
(cd ab_r1 && timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > ../gpurun_out/r2c/r1_cfg4.json 2> ../gpurun_out/r2c/r1_cfg4.err)
tail -c 300 gpurun_out/r2c/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],3),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',r.get('frac'),'share',{k:round(v,2) for k,v in r['share_of_step'].items()})
    except Exception as e: print(f,'ERR',e)
PY
