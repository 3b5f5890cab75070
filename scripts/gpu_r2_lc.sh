#!/bin/bash
# cached work lists of resident plans: tests, then A/B (TB_NO_LIST_CACHE=1) of the resident step on cfg5 / cfg3 / cfg2 / cfg4-8
O=gpurun_out/r2lc; mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests/test_list_cache.py -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_list_cache.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_all.txt
B="--no-cpu-baseline --no-other-configs --steps 30 --warmup 5"
for wl in cfg5 cfg3 cfg2; do
  timeout 600 python bench.py --workload $wl $B > $O/cache_$wl.json 2> $O/cache_$wl.err
  TB_NO_LIST_CACHE=1 timeout 600 python bench.py --workload $wl $B > $O/nocache_$wl.json 2> $O/nocache_$wl.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2lc/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d.get('roofline',{})
        print(f.split('/')[-1], 'ms',round(d['ms_per_step'],4),'median',round(d.get('ms_per_step_median_rank0',0),4),'Gop/s',round(d['value']),'launches',d.get('launches_per_step'),'frac',r.get('frac'),'whole',r.get('frac_whole_step'),'hybrid',r.get('hybrid',{}).get('frac_whole_step'),'e2e ms',d.get('e2e',{}).get('ms_per_step'),d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
for f in $O/*.err; do echo "== $f"; tail -c 300 $f; done
