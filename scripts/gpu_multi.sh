#!/bin/bash
# Multi-GPU check (run with: gpurun --gpus N -- bash scripts/gpu_multi.sh TAG N [workload])
TAG=${1:-m01}
N=${2:-2}
WL=${3:-cfg2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpus.csv 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== 1 GPU"; timeout 600 python bench.py --gpus 1 --workload $WL --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n1.json
echo "== $N GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $WL --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | tail -3 | tee $OUT/bench_n$N.json
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --workload $WL --steps 1 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_ref_n$N.json
ls -la $OUT
