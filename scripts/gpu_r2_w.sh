#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O; rm -f $O/*
./tensorbranching.jl_b200/dpx_peak > $O/dpx_peak.json 2>&1; cat $O/dpx_peak.json
timeout 600 python bench.py --workload cfg2 --weights f32 --steps 10 --warmup 3 --no-other-configs > $O/bench_cfg2_f32.json 2> $O/bench_cfg2_f32.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2w/bench_cfg2_f32.json').read().strip().splitlines() if l.startswith('{')][-1])
r=d['roofline']; print('ms',d['ms_per_step'],'Gop/s',d['value'],'peak',r['peak'],'frac',r['frac'],r['frac_single_lane'],r['frac_whole_step'],'e2e',d['e2e']['vs_resident'])
PY
