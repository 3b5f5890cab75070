#!/bin/bash
# all-optimal-configuration tables: tests, timing, memcheck of the new kernels
O=gpurun_out/r2z; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_table_configs.py tests/test_wide_value_types.py -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_table.txt
timeout 300 python scripts/table/time_table_configs.py 2>&1 | tee $O/table_configs_timing.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_table_configs.py -m gpu -x -q -k "test_gpu_table_configs and not large" 2>&1 | tail -8 | tee $O/memcheck_table.txt
