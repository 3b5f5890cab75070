#!/bin/bash
# consumer-warp phase accounting of the persistent GEMM kernel (library built with -DTB_KPROF) on the dataflow workloads
O=gpurun_out/r2t; mkdir -p $O; rm -f $O/*
export TBCUDA_LIB=$PWD/tensorbranching.jl_b200/libtbcuda_kprof.so
for w in cfg5 cfg3 cfg2; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-other-configs --no-e2e --cpu-budget 0 2>&1 | grep -E "KPROF|ms_per_step" | cut -c1-300 > $O/kprof_$w.txt
  TB_TL_DUMP=$O/tl_$w.bin timeout 300 python scripts/diag/timeline.py $w > $O/tl_run_$w.log 2>&1
  python scripts/diag/timeline.py --analyze $O/tl_$w.bin > $O/tl_analysis_$w.txt 2>&1
  rm -f $O/tl_$w.bin
done
for w in cfg5 cfg3 cfg2; do echo "== $w"; grep KPROF $O/kprof_$w.txt | tail -1; cat $O/tl_analysis_$w.txt; done
