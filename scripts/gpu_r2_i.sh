#!/bin/bash
# single-barrier epilogue of k_gemm2h: cfg2 / cfg4(8) / cfg5 against the previous build (ab_prev/) on one box + parity tests
O=gpurun_out/r2i; mkdir -p $O; rm -f $O/*
B="--no-cpu-baseline --no-e2e --no-other-configs"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dataflow.py tests/test_baseline_configs.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest.txt; cat $O/pytest.txt
for rep in 1 2; do
  timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 $B > $O/new_cfg2_$rep.json 2> $O/new_cfg2_$rep.err
  TBCUDA_LIB=$PWD/ab_prev/libtbcuda.so timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 $B > $O/prev_cfg2_$rep.json 2> $O/prev_cfg2_$rep.err
done
timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 $B > $O/new_cfg4.json 2> $O/new_cfg4.err
TBCUDA_LIB=$PWD/ab_prev/libtbcuda.so timeout 300 python bench.py --workload cfg4 --max-branches 8 --steps 3 --warmup 3 $B > $O/prev_cfg4.json 2> $O/prev_cfg4.err
timeout 300 python bench.py --workload cfg5 --steps 100 --warmup 20 $B > $O/new_cfg5.json 2> $O/new_cfg5.err
TBCUDA_LIB=$PWD/ab_prev/libtbcuda.so timeout 300 python bench.py --workload cfg5 --steps 100 --warmup 20 $B > $O/prev_cfg5.json 2> $O/prev_cfg5.err
tail -c 300 $O/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d['roofline']
        print(f.split('/')[-1],'ms',round(d['ms_per_step'],4),'median',round(d['ms_per_step_median_rank0'],4),'Gop/s',round(d['value']),'frac',round(r.get('frac') or 0,3),'1lane',round(r.get('frac_single_lane') or 0,3),'share',{k:round(v,2) for k,v in r['share_of_step'].items()},d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
