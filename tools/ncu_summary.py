"""Key metrics of one ncu report (raw page) per element: python tools/ncu_summary.py REP N_ELEM"""
import csv
import subprocess
import sys

rep, nel = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
m = {n: (v[i], u[i]) for i, n in enumerate(h)}
def f(n):
    try:
        return float(m[n][0].replace(",", ""))
    except Exception:
        return float("nan")
print("time ms", f("gpu__time_duration.sum") if m["gpu__time_duration.sum"][1] == "ms" else m["gpu__time_duration.sum"])
print("regs", m["launch__registers_per_thread"][0])
print("inst/elem", f("smsp__inst_executed.sum") / nel)
print("smem wavefronts/elem", f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / nel, "conflicts", f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / nel)
print("st sectors/elem", f("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum") / nel, "st requests/elem", f("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum") / nel)
print("cycles/elem/SM", f("sm__cycles_active.avg") * 148 / nel)
for n in ["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum", "dram__bytes_read.sum",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]:
    if n in m:
        print(n, m[n])
st = {n: f(n) for n in h if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")}
for n, x in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
    print("  stall", n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), round(x, 2))
