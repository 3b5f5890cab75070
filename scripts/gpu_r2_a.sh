#!/bin/bash
# round 2, first validation of the dataflow executor: GPU tests, then cfg2/cfg3/cfg5 with both executors
set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a/pytest.txt; cat gpurun_out/r2a/pytest.txt
for wl in cfg2 cfg3 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-budget 3 > gpurun_out/r2a/bench_${wl}_dataflow.json 2> gpurun_out/r2a/bench_${wl}_dataflow.err
  TB_LEVEL_SYNC=1 timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench_${wl}_levelsync.json 2> gpurun_out/r2a/bench_${wl}_levelsync.err
done
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --max-branches 8 > gpurun_out/r2a/bench_cfg4_8_dataflow.json 2> gpurun_out/r2a/bench_cfg4_8_dataflow.err
tail -c 600 gpurun_out/r2a/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],3),'Gop/s',round(d['value']),'launches/step',d['launches_per_step'],'frac',d['roofline']['frac'],'e2e',round(d['e2e']['value']) if 'e2e' in d else None, d.get('cpu_baseline',{}).get('agrees_with_gpu'))
    except Exception as e: print(f, 'ERR', e)
PY
