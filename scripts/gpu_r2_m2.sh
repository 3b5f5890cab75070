#!/bin/bash
# 2 GPUs: the in-library multi-GPU context over real NCCL (tests + plain-C client), and bench.py strong scaling
mkdir -p gpurun_out/r2m2; rm -f gpurun_out/r2m2/*
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu_device.py tests/test_abi_layout.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2m2/pytest.txt; cat gpurun_out/r2m2/pytest.txt
B="--no-cpu-baseline --no-other-configs"
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 $B > gpurun_out/r2m2/n1_cfg4.json 2> gpurun_out/r2m2/n1_cfg4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 $B > gpurun_out/r2m2/n2_cfg4.json 2> gpurun_out/r2m2/n2_cfg4.err
timeout 600 python bench.py --gpus 1 --workload cfg2 --steps 10 --warmup 3 $B > gpurun_out/r2m2/n1_cfg2.json 2> gpurun_out/r2m2/n1_cfg2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg2 --steps 10 --warmup 3 $B > gpurun_out/r2m2/n2_cfg2.json 2> gpurun_out/r2m2/n2_cfg2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg3 --steps 10 --warmup 3 $B > gpurun_out/r2m2/n2_cfg3.json 2> gpurun_out/r2m2/n2_cfg3.err
tail -c 600 gpurun_out/r2m2/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m2/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d['roofline']
        print(f.split('/')[-1], 'N',d['n_gpus'],d['scaling'],'ms',round(d['ms_per_step'],3),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',round(r.get('frac') or 0,3),'e2e',round(d['e2e']['value']),'e2e ms',round(d['e2e']['ms_per_step'],3),d['e2e']['host_breakdown_rank0'].get('compile_ms'),d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
