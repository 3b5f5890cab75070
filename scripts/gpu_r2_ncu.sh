#!/bin/bash
# round 2 ncu evidence: launch lists of the default bench command and of cfg2 / cfg3 / cfg5, full captures of the
# persistent kernel with the engine's per-launch algorithmic bytes of THE SAME launches (cfg4 dominant launch, cfg2,
# cfg3 dataflow instance, f32-weighted cfg2s)
O=gpurun_out/r2ncu; mkdir -p $O; rm -f $O/*
NCU="ncu --clock-control none"
B="--no-cpu-baseline --no-e2e --no-other-configs"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 $B > $O/launches_default.log 2>&1
for wl in cfg2 cfg3 cfg5; do
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file $O/launches_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 $B > $O/launches_$wl.log 2>&1
done
timeout 900 $NCU --set full --import-source on -k regex:k_gemm2 --launch-skip 4 -c 1 -f -o $O/full_cfg4 python scripts/ncu_traffic.py --workload cfg4 --count 4 --dump $O/dump_cfg4.csv > $O/full_cfg4.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_gemm2 -c 4 -f -o $O/full_cfg2 python scripts/ncu_traffic.py --workload cfg2 --count 512 --dump $O/dump_cfg2.csv > $O/full_cfg2.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_gemm2 -c 1 -f -o $O/full_cfg3 python scripts/ncu_traffic.py --workload cfg3 --dump $O/dump_cfg3.csv > $O/full_cfg3.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_gemm2 -c 3 -f -o $O/full_cfg2s_f32 python scripts/ncu_traffic.py --workload cfg2s --count 256 --weights f32 --dump $O/dump_cfg2s_f32.csv > $O/full_cfg2s_f32.log 2>&1
for f in $O/*.log; do echo "== $f"; tail -n 3 $f; done; ls -la $O
