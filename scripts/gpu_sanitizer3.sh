#!/bin/bash
# racecheck over the kernels that do NOT synchronise through mbarriers (fused subtrees, generic, cp.async GEMM v1, permute):
# int32 / f32 plans with TB_GEMM_V1=1
TAG=${1:-z03}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
TB_GEMM_V1=1 timeout 200 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(kernel_paths and (64-100-7 or 72-100-7 or 66-100-7 or 80-100-7 or 64-120-9)) or weighted_f32 or permute_bits" > $OUT/racecheck_pytest.log 2>&1
echo "racecheck exit $?" | tee $OUT/racecheck.rc
tail -3 $OUT/racecheck_pytest.log; tail -4 $OUT/racecheck.log
