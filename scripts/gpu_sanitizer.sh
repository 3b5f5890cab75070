#!/bin/bash
# compute-sanitizer evidence (bounded): memcheck over the all-kernel-paths parity tests of one graph size + the golden
# fixtures, racecheck over smoke().  Summaries go to gpurun_out/<tag>/ (copy into profiles/).
TAG=${1:-z01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== memcheck"
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(single_plan_all_kernel_paths and 100-7) or golden or packed_int16" > $OUT/memcheck_pytest.log 2>&1
echo "memcheck exit $?" | tee $OUT/memcheck.rc
tail -3 $OUT/memcheck_pytest.log; tail -4 $OUT/memcheck.log
echo "== racecheck"
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file $OUT/racecheck.log \
    python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck_smoke.log 2>&1
echo "racecheck exit $?" | tee $OUT/racecheck.rc
tail -2 $OUT/racecheck_smoke.log; tail -4 $OUT/racecheck.log
ls -la $OUT
