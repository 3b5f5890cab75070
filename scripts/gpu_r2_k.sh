#!/bin/bash
# end-to-end (host buffers, plans compiled inside the timed region) after the plan-compiler speed-up
O=gpurun_out/r2k; mkdir -p $O; rm -f $O/*
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg2.json 2> $O/cfg2.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg5.json 2> $O/cfg5.err
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1],'ms',round(d['ms_per_step'],4),'e2e',d['e2e']['ms_per_step'],d['e2e']['steps'],d['e2e']['host_breakdown_rank0'])
    except Exception as e: print(f,'ERR',e)
PY
