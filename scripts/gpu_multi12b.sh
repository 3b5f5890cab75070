#!/bin/bash
# 8-GPU check of the driver's default command (weak scaling) + cfg2 strong for the record
TAG=${1:-m12b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== cfg2 N=8 weak (driver command)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_cfg2_n8_weak.out 2> $OUT/bench_cfg2_n8_weak.err; echo "exit $?" | tee $OUT/weak.rc
tail -c 400 $OUT/bench_cfg2_n8_weak.out
echo "== cfg2 N=8 strong"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 8 --steps 20 --warmup 5 --scaling strong --no-cpu-baseline > $OUT/bench_cfg2_n8_strong.out 2> $OUT/bench_cfg2_n8_strong.err; echo "exit $?" | tee $OUT/strong.rc
tail -c 300 $OUT/bench_cfg2_n8_strong.out
ls -la $OUT
