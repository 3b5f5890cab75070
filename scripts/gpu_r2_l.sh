#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O; rm -f $O/*
timeout 300 python scripts/trace_call.py cfg2 30 > $O/trace_cfg2.txt 2>&1
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg2.json 2> $O/cfg2.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg5.json 2> $O/cfg5.err
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
grep "call\|stall" $O/trace_cfg2.txt | grep -v "call [0-9]*$" | tail -40
grep -o '"e2e": {[^}]*}' $O/cfg2.json | cut -c1-900
grep -o '"e2e": {[^}]*}' $O/cfg5.json | cut -c1-900
