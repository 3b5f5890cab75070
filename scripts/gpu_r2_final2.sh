#!/bin/bash
# end of round 2, after the cached work lists: all GPU tests, the driver's default command, f32 line, cfg1, smoke, ncu launch list of cfg5
O=gpurun_out/r2final2; mkdir -p $O; rm -f $O/*
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest.txt; cat $O/pytest.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --workload cfg2 --weights f32 --steps 10 --warmup 3 --no-other-configs > $O/bench_cfg2_f32.json 2> $O/bench_cfg2_f32.err
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-other-configs > $O/bench_cfg1.json 2> $O/bench_cfg1.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg5.csv python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > $O/launches_cfg5.log 2>&1
for f in $O/*.err; do echo "== $f"; tail -c 300 $f; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2final2/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d.get('roofline',{})
        print(f.split('/')[-1], d.get('impl'), 'ms',round(d['ms_per_step'],3),'Gop/s',round(d['value']),'launches',d.get('launches_per_step'),'frac',r.get('frac'),'1lane',r.get('frac_single_lane'),'whole',r.get('frac_whole_step'),'e2e',d.get('e2e',{}).get('value'),d.get('e2e',{}).get('vs_resident'),d.get('agrees_with_golden'))
        for k,v in d.get('other_configs',{}).items(): print('   ',k,{q:(round(x,3) if isinstance(x,float) else x) for q,x in v.items() if q not in('workload','e2e')}, 'e2e', round(v['e2e']['ms_per_step'],3), round(v['e2e']['vs_resident'],3))
    except Exception as e: print(f,'ERR',e)
PY
