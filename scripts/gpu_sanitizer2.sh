#!/bin/bash
# broader compute-sanitizer pass: memcheck over the GPU suite without the multi-GB cases, racecheck over the fixtures
TAG=${1:-z02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== memcheck"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/memcheck.log \
    python -m pytest tests -m gpu -q -k "not sc28 and not sc24 and not full_size and not large and not config2" > $OUT/memcheck_pytest.log 2>&1
echo "memcheck exit $?" | tee $OUT/memcheck.rc
tail -3 $OUT/memcheck_pytest.log; tail -4 $OUT/memcheck.log
echo "== racecheck"
timeout 160 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file $OUT/racecheck.log \
    python -m pytest tests -m gpu -q -k "golden or packed_int16 or sliced_matches or contract_tensor or stream_equals" > $OUT/racecheck_pytest.log 2>&1
echo "racecheck exit $?" | tee $OUT/racecheck.rc
tail -3 $OUT/racecheck_pytest.log; tail -4 $OUT/racecheck.log
