#!/bin/bash
TAG=${1:-t01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TBCUDA_LIB=$PWD/tensorbranching.jl_b200/libtbcuda_kprof.so TB_TL_DUMP=$OUT/tl.bin timeout 300 python scripts/diag/timeline.py cfg2 2>&1 | tail -8 | tee $OUT/run.log
python scripts/diag/timeline.py --analyze $OUT/tl.bin | tee $OUT/analysis.txt
ls -la $OUT
