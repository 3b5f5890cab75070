#!/bin/bash
# 2-GPU strong-scaling points of the sharded configs (completes the 1/2/4/8 series of cfg4)
TAG=${1:-m12c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
run() { # N workload steps warmup extra
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $1 --workload $2 --steps $3 --warmup $4 --no-cpu-baseline --scaling strong $5 2>&1 | tail -1 | tee $OUT/bench_$2_n$1.json
}
echo "== cfg4 N=2"; run 2 cfg4 3 3
echo "== cfg5 N=2"; run 2 cfg5 20 5
echo "== cfg3 N=2"; run 2 cfg3 20 5
ls -la $OUT
