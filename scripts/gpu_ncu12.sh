#!/bin/bash
# round-12 ncu evidence for the default command (cfg2): launch list (time only) + one --set full capture of k_gemm2h
TAG=${1:-n12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== ncu launches"
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
echo "exit $?"; tail -c 200 $OUT/ncu_launches.log
echo "== ncu full (gemm)"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_gemm2h -s 300 -c 4 -o $OUT/prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
echo "exit $?"
ls -la $OUT
