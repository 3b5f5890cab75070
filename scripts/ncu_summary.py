"""Summarise an `ncu --set full` capture of the dominant GEMM kernel into profiles/:
   <tag>_ncu_full_<kernel>_summary.csv   selected raw metrics per captured launch
   <tag>_ncu_traffic.json                DRAM bytes per launch (roofline.traffic of bench.py)
usage: python scripts/ncu_summary.py gpurun_out/r08/prof_gemm.ncu-rep r08"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    u = unit.lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kname = data[0][hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("tb::", "").replace("<", "_").replace(">", "")
    keep = [h for h in hdr if h in KEEP or "average_warps_issue_stalled" in h]
    with open(f"profiles/{tag}_ncu_full_{kname}_summary.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i + 1}" for i in range(len(data))])
        for k in keep:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in data])
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    launches = [{"dram_bytes": to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]), "duration_us": float(r[it].replace(",", "")),
                 "grid": int(r[hdr.index("launch__grid_size")])} for r in data]
    out = {"kernel": kname, "source": rep, "launches": launches,
           "dram_bytes_per_launch_mean": sum(l["dram_bytes"] for l in launches) / len(launches)}
    json.dump(out, open(f"profiles/{tag}_ncu_traffic.json", "w"), indent=1)
    json.dump(out, open("profiles/latest_ncu_traffic.json", "w"), indent=1)
    print(json.dumps(out)[:400])


if __name__ == "__main__":
    main()
