#!/bin/bash
# fused-subtree launches chained across waves (so that a wave's big-step kernel starts beside the next wave's fused launch)
O=gpurun_out/r2v; mkdir -p $O; rm -f $O/*
run() { name=$1; shift
  for w in cfg5 cfg2 cfg3; do
    env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-other-configs --no-e2e --cpu-budget 0 > $O/${w}_$name.json 2> $O/${w}_$name.err
  done
}
run base TB_NOOP=1
run chain TB_FUSED_CHAIN=1
run chain_c100 TB_FUSED_CHAIN=1 TB_FUSED_CARVEOUT=100
run base2 TB_NOOP=1
run chain2 TB_FUSED_CHAIN=1
run chain_c100b TB_FUSED_CHAIN=1 TB_FUSED_CARVEOUT=100
TB_FUSED_CHAIN=1 TB_FUSED_CARVEOUT=100 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py tests/test_dataflow.py -m gpu -x -q 2>&1 | tail -2
tail -c 200 $O/*.err | tail -6
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2v/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline'].get('frac') or 0,3), d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
