#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
timeout 300 python scripts/trace_call.py cfg2 8 > $O/trace_cfg2.txt 2>&1
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg2.json 2> $O/cfg2.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg5.json 2> $O/cfg5.err
timeout 300 python scripts/permute_bw.py > $O/permute_bw.txt 2>&1
grep "call\|stall" $O/trace_cfg2.txt | grep -v "call [0-9]*$" | tail -8
for w in cfg2 cfg5; do grep -o '"ms_per_step": [0-9.]*' $O/$w.json | head -1; grep -o '"e2e": {[^}]*}' $O/$w.json | cut -c1-700; done
cat $O/permute_bw.txt
