#!/bin/bash
# 2 GPUs after the cached work lists: in-library multi-GPU tests over real NCCL, bench.py strong scaling (cfg4, cfg5)
O=gpurun_out/r2m2b; mkdir -p $O; rm -f $O/*
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu_device.py tests/test_list_cache.py -m gpu -x -q 2>&1 | tail -8 > $O/pytest.txt; cat $O/pytest.txt
B="--no-cpu-baseline --no-other-configs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 $B > $O/n2_cfg4.json 2> $O/n2_cfg4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5 --steps 20 --warmup 3 $B > $O/n2_cfg5.json 2> $O/n2_cfg5.err
tail -c 400 $O/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m2b/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        r=d['roofline']
        print(f.split('/')[-1], 'N',d['n_gpus'],d['scaling'],'ms',round(d['ms_per_step'],3),'Gop/s',round(d['value']),'launches',d['launches_per_step'],'frac',round(r.get('frac') or 0,3),'e2e',round(d['e2e']['value']),'e2e ms',round(d['e2e']['ms_per_step'],3),d.get('agrees_with_golden'))
    except Exception as e: print(f,'ERR',e)
PY
