#!/bin/bash
# validation call (gpurun --gpus 2 -- bash scripts/gpu_val12.sh TAG): full GPU test-suite, default bench, 2-rank weak and strong
TAG=${1:-v12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_cfg2_n1.json
run() { # N scaling
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $1 --no-cpu-baseline --scaling $2 2>&1 | tail -1 | tee $OUT/bench_cfg2_n$1_$2.json
}
echo "== cfg2 N=2 weak"; run 2 weak
echo "== cfg2 N=2 strong"; run 2 strong
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_n2.json
ls -la $OUT
