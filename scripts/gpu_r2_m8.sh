#!/bin/bash
# 8 GPUs: the driver's scaling command (strong, cfg4), cfg2 strong, and the in-library multi-GPU context over 8 devices
O=gpurun_out/r2m8; mkdir -p $O; rm -f $O/*
nvidia-smi -L | wc -l
B="--no-cpu-baseline --no-other-configs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8_cfg4.json 2> $O/n8_cfg4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > $O/n4_cfg4.json 2> $O/n4_cfg4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --workload cfg2 --steps 20 --warmup 5 $B > $O/n8_cfg2.json 2> $O/n8_cfg2.err
timeout 600 python -m pytest tests/test_multi_gpu_device.py tests/test_abi_layout.py -m gpu -x -q 2>&1 | tail -6 > $O/pytest.txt; cat $O/pytest.txt
for f in $O/*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $f | tail -c 500; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m8/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d['roofline']
        print(f.split('/')[-1], 'N',d['n_gpus'],d['scaling'],'ms',round(d['ms_per_step'],3),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',round(r.get('frac') or 0,3),'e2e',round(d['e2e']['value']),'e2e ms',round(d['e2e']['ms_per_step'],3),d['e2e']['host_breakdown_rank0'],d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
