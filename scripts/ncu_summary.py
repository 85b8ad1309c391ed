"""Summarise an .ncu-rep (raw page CSV) into a small per-kernel table for profiles/."""
import csv, subprocess, sys, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(r, k):
    return r[hdr.index(k)] if k in hdr else ""
keys = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr_per_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("lts__t_sectors_op_red.sum", "l2_red_sectors"), ("lts__t_sectors_op_atom.sum", "l2_atom_sectors"),
        ("l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "red_ops")]
out = []
for r in rows[2:]:
    d = {"kernel": col(r, "Kernel Name").split("(")[0][:48], "grid": col(r, "Grid Size"), "block": col(r, "Block Size")}
    for k, n in keys:
        v = col(r, k)
        if v not in ("", "n/a"):
            u = units[hdr.index(k)]
            d[n] = f"{v} {u}".strip()
    st = [(hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
          for i in range(len(hdr)) if "issue_stalled" in hdr[i] and hdr[i].endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")]
    st.sort(key=lambda x: -x[1])
    d["top_stalls"] = ", ".join(f"{n}={v:.2f}" for n, v in st[:4])
    out.append(d)
for d in out:
    print(json.dumps(d))
