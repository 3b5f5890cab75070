#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O; rm -f $O/*
timeout 300 python scripts/plan_bench/run.py cfg2 > $O/plan_bench_cfg2.txt 2>&1
timeout 300 python scripts/trace_call.py cfg2 12 > $O/trace_cfg2.txt 2>&1
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg2.json 2> $O/cfg2.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg5.json 2> $O/cfg5.err
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg3.json 2> $O/cfg3.err
cat $O/plan_bench_cfg2.txt
grep "call\|stall" $O/trace_cfg2.txt | grep -v "call [0-9]*$" | tail -14
for w in cfg2 cfg5 cfg3; do grep -o '"e2e": {[^}]*}' $O/$w.json | cut -c1-700; done
