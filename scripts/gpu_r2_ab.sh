#!/bin/bash
# A/B: A operand of k_gemm2h through LDS.128 when mp == 1 (in-tree build) against the previous build (ab_head/)
O=gpurun_out/r2ab; mkdir -p $O; rm -f $O/*
for rep in 1 2; do
for v in new head; do
  if [ $v = head ]; then export TBCUDA_LIB=$PWD/ab_head/libtbcuda.so; else unset TBCUDA_LIB; fi
  timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${v}_cfg2_$rep.json 2> $O/${v}_cfg2_$rep.err
  timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${v}_cfg5_$rep.json 2> $O/${v}_cfg5_$rep.err
  timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${v}_cfg3_$rep.json 2> $O/${v}_cfg3_$rep.err
done
done
for v in new head; do
  if [ $v = head ]; then export TBCUDA_LIB=$PWD/ab_head/libtbcuda.so; else unset TBCUDA_LIB; fi
  timeout 600 python bench.py --workload cfg4 --max-branches 8 --steps 5 --warmup 3 --no-other-configs --no-e2e --cpu-budget 0 > $O/${v}_cfg4.json 2> $O/${v}_cfg4.err
done
unset TBCUDA_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -2
tail -c 300 $O/*.err | tail -20
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2ab/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'frac', d['roofline'].get('frac'), d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
