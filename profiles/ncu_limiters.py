"""Per-kernel limiter table of an ncu --set full report: duration, throughput of the memory levels and pipes in % of peak, occupancy, and the
two largest warp-stall reasons (stall cycles per issued instruction).  usage: python profiles/ncu_limiters.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return float("nan")
M = {"us": "gpu__time_duration.sum", "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "L2%": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "L1%": "l1tex__throughput.avg.pct_of_peak_sustained_active", "issue%": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "fp64%": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "alu%": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "lsu%": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "warps%": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "regs": "launch__registers_per_thread", "Minst": "smsp__inst_executed.sum"}
stalls = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("per_warp_active.pct")] or \
         [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
print(f"{'kernel':30s} " + " ".join(f"{k:>7s}" for k in M) + "  top stalls")
for r in rows[2:]:
    n = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")[:30]
    vals = []
    for k, m in M.items():
        c = col(m)
        v = f(r[c]) if c is not None else float("nan")
        if k == "us": v *= scale.get(units[c], 1.0)
        if k == "Minst": v /= 1e6
        vals.append(v)
    st = sorted(((f(r[hdr.index(h)]), h.split("issue_stalled_")[1].split("_per_")[0]) for h in stalls), reverse=True)[:3]
    print(f"{n:30s} " + " ".join(f"{v:7.1f}" for v in vals) + "  " + ", ".join(f"{b} {a:.2f}" for a, b in st))
