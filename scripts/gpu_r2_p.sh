#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg2.json 2> $O/cfg2.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-other-configs --cpu-budget 3 > $O/cfg5.json 2> $O/cfg5.err
for w in cfg2 cfg5; do grep -o '"ms_per_step": [0-9.]*' $O/$w.json | head -1; grep -o '"e2e": {[^}]*}' $O/$w.json | cut -c1-700; done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_wide_value_types.py "tests/test_gpu_parity.py::test_permute_bits" -m gpu -x -q > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee $O/memcheck.rc
tail -5 $O/memcheck.log
