#!/bin/bash
# 8-GPU call of round 12 (gpurun --gpus 8 -- bash scripts/gpu_multi12.sh TAG): stream tests on one GPU, then the
# sharded configs at N = 8 (and cfg4 at N = 4).
TAG=${1:-m12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpus.csv 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== stream + slicing tests"; timeout 300 python -m pytest tests/test_stream.py tests/test_index_slicing.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_stream.log
run() { # N workload steps warmup extra
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $1 --workload $2 --steps $3 --warmup $4 --no-cpu-baseline --scaling strong $5 2>&1 | tail -1 | tee $OUT/bench_$2_n$1.json
}
echo "== cfg4 N=8"; run 8 cfg4 3 3
echo "== cfg4 N=4"; run 4 cfg4 3 3 --no-e2e
echo "== cfg3 N=8 (index sliced)"; run 8 cfg3 20 5
echo "== cfg5 N=8"; run 8 cfg5 20 5
ls -la $OUT
