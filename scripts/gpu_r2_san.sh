#!/bin/bash
# compute-sanitizer memcheck over the round-2 device code: dataflow instances (forced), 8-byte value types, multi-device
# context on a repeated device, solo waves; then racecheck over one forced-dataflow contraction (shared-memory hazards
# between the mbarrier-ordered producer / consumer warps are expected, see profiles/z02_sanitizer_racecheck_summary.csv)
O=gpurun_out/r2san; mkdir -p $O; rm -f $O/*
echo "== memcheck (dataflow forced)"
TB_DATAFLOW=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $O/memcheck_dataflow.log \
    python -m pytest tests/test_gpu_parity.py tests/test_dataflow.py -m gpu -x -q -k "golden or (single_plan_all_kernel_paths and 100-7) or many_level or packed_int16 or solo or f32_with_gemm" > $O/memcheck_dataflow_pytest.log 2>&1
echo "memcheck dataflow exit $?" | tee $O/memcheck_dataflow.rc
tail -3 $O/memcheck_dataflow_pytest.log; tail -3 $O/memcheck_dataflow.log
echo "== memcheck (8-byte value types, multi-device context)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $O/memcheck_wide_multi.log \
    python -m pytest tests/test_wide_value_types.py tests/test_multi_gpu_device.py tests/test_golden_sliced_open.py -m gpu -x -q > $O/memcheck_wide_multi_pytest.log 2>&1
echo "memcheck wide/multi exit $?" | tee $O/memcheck_wide_multi.rc
tail -3 $O/memcheck_wide_multi_pytest.log; tail -3 $O/memcheck_wide_multi.log
echo "== synccheck (dataflow forced)"
TB_DATAFLOW=1 timeout 600 compute-sanitizer --tool synccheck --error-exitcode 77 --log-file $O/synccheck_dataflow.log \
    python -m pytest tests/test_dataflow.py -m gpu -x -q -k "golden" > $O/synccheck_dataflow_pytest.log 2>&1
echo "synccheck exit $?" | tee $O/synccheck.rc
tail -2 $O/synccheck_dataflow_pytest.log; tail -3 $O/synccheck_dataflow.log
ls -la $O
