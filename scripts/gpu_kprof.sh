#!/bin/bash
# Diagnostics: SM-cycle accounting of the packed GEMM consumers (library built with -DTB_KPROF).
TAG=${1:-k01}; WL=${2:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for combo in "4 128" "1 128" "8 128"; do
  set -- $combo
  echo "== lanes $1 wave $2" | tee -a $OUT/kprof.log
  TBCUDA_LIB=$PWD/tensorbranching.jl_b200/libtbcuda_kprof.so TB_LANES=$1 TB_WAVE=$2 timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "KPROF|ms_per_step" | cut -c1-400 | tee -a $OUT/kprof.log
done
