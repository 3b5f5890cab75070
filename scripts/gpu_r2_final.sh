#!/bin/bash
# end of round 2: all GPU tests, the driver's two commands exactly (reference arm, default arm), f32 line, cfg1, smoke,
# the ncu launch list of the default command and one full capture of the branching-table kernel
O=gpurun_out/r2final; mkdir -p $O; rm -f $O/*
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest.txt; cat $O/pytest.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --workload cfg2 --weights f32 --steps 10 --warmup 3 --no-other-configs > $O/bench_cfg2_f32.json 2> $O/bench_cfg2_f32.err
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-other-configs > $O/bench_cfg1.json 2> $O/bench_cfg1.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > $O/launches_default.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:k_region_configs --launch-skip 60 -c 3 -f -o $O/full_region_configs python scripts/table/time_table_configs.py > $O/full_region_configs.log 2>&1
for f in $O/*.err; do echo "== $f"; tail -c 300 $f; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2final/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d.get('roofline',{})
        print(f.split('/')[-1], d.get('impl'), 'ms',round(d['ms_per_step'],3),'Gop/s',round(d['value']),'launches',d.get('launches_per_step'),'frac',r.get('frac'),'1lane',r.get('frac_single_lane'),'whole',r.get('frac_whole_step'),'e2e',d.get('e2e',{}).get('value'),d.get('e2e',{}).get('vs_resident'),d.get('agrees_with_golden'),d.get('cpu_baseline'))
        for k,v in d.get('other_configs',{}).items(): print('   ',k,{q:(round(x,3) if isinstance(x,float) else x) for q,x in v.items() if q!='workload'})
    except Exception as e: print(f,'ERR',e)
PY
