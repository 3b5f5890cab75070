#!/bin/bash
# 8-GPU box: strong-scaling points N=8 and N=4 of the resident + e2e bench
TAG=${1:-m08}; WL=${2:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.csv 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for N in 8 4; do
  echo "== $N GPUs"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --workload $WL --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | tail -2 | tee $OUT/bench_n$N.json
done
ls -la $OUT
