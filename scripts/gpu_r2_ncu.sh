#!/bin/bash
# round 2 ncu evidence: launch list of the default bench command, full captures of the persistent kernel with the
# engine's per-launch algorithmic bytes of the same launches (cfg4, cfg2, f32-weighted cfg2s)
O=gpurun_out/r2ncu; mkdir -p $O
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > $O/launches_default.log 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_cfg2.csv python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > $O/launches_cfg2.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_gemm2 -c 1 -f -o $O/full_cfg4 python scripts/ncu_traffic.py --workload cfg4 --count 4 --dump $O/dump_cfg4.csv > $O/full_cfg4.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_gemm2 -c 3 -f -o $O/full_cfg2 python scripts/ncu_traffic.py --workload cfg2 --count 512 --dump $O/dump_cfg2.csv > $O/full_cfg2.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_gemm2 -c 2 -f -o $O/full_cfg2s_f32 python scripts/ncu_traffic.py --workload cfg2s --count 256 --weights f32 --dump $O/dump_cfg2s_f32.csv > $O/full_cfg2s_f32.log 2>&1
tail -3 $O/*.log; ls -la $O
