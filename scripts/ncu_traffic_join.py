#!/usr/bin/env python
"""Join an `ncu --set full` capture of the persistent GEMM kernel with the engine's per-launch records of the same run
(scripts/ncu_traffic.py): per captured launch the DRAM bytes ncu measured and the algorithmic bytes / ops the launch
stands for.  usage: ncu_traffic_join.py <rep.ncu-rep> <dump.csv> <out.json> <workload note>"""
import csv
import io
import json
import subprocess
import sys

rep, dump, out, note = sys.argv[1:5]
skip = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # launches of the kernel skipped by ncu (--launch-skip)
METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "sm__inst_executed_pipe_alu.sum", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
units = rows[1]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3, "nsecond": 1e-6}.get(unit, 1)


launches = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    rec = {"kernel": r[col["Kernel Name"]][:60]}
    rec["dram_read"] = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    rec["dram_write"] = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    rec["ncu_ms"] = to_ms(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
    for m in METRICS[3:]:
        if m in col:
            try:
                rec[m] = float(r[col[m]].replace(",", ""))
            except ValueError:
                pass
    launches.append(rec)
eng = []
for ln in open(dump):
    f = ln.strip().split(",")
    if len(f) >= 7 and int(f[0]) == 2:
        eng.append(dict(event_ms=float(f[2]) - float(f[1]), ops=float(f[3]), algorithmic_bytes=float(f[4]), instances=int(f[5]), tiles=int(f[6])))
eng = eng[skip:]
n = min(len(launches), len(eng))
joined = []
for a, b in zip(launches[:n], eng[:n]):
    d = dict(a)
    d.update(b)
    d["dram_bytes"] = a["dram_read"] + a["dram_write"]
    d["traffic_over_algorithmic"] = d["dram_bytes"] / b["algorithmic_bytes"] if b["algorithmic_bytes"] else None
    joined.append(d)
res = {"note": note, "launches": joined,
       "dram_bytes_per_launch_mean": sum(d["dram_bytes"] for d in joined) / max(1, n),
       "algorithmic_bytes_per_launch_mean": sum(d["algorithmic_bytes"] for d in joined) / max(1, n),
       "traffic_over_algorithmic": (sum(d["dram_bytes"] for d in joined) / max(1.0, sum(d["algorithmic_bytes"] for d in joined))) if n else None}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "launches"}, indent=1))
for d in joined:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items()})
