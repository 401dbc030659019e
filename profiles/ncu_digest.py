"""Digest of an ncu report for one kernel: headline metrics, dynamic instruction mix, stall mix and warp-instructions
per barrier-delimited phase.  usage: python profiles/ncu_digest.py report.ncu-rep [kernel-index]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2 + kidx]
print("kernel:", r[hdr.index("Kernel Name")].split("(")[0])
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_shared_mem",
          "launch__occupancy_limit_registers", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]:
    if k in hdr:
        print(f"  {k:70s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
lo = his[kidx]; hi = his[kidx + 1] - 1 if kidx + 1 < len(his) else None
hdr = rows[lo]; ix = {h: i for i, h in enumerate(hdr)}
data = [x for x in rows[lo + 1:hi] if len(x) > 10]
tot = sum(float(x[ix["Instructions Executed"]] or 0) for x in data)
agg = collections.Counter()
for x in data:
    op = [y for y in x[ix["Source"]].split() if not y.startswith("@")]
    agg[op[0].split(".")[0] if op else ""] += float(x[ix["Instructions Executed"]] or 0)
print("instruction mix (warp-level):")
for k, v in agg.most_common(14):
    print(f"  {k:10s} {v/1e6:8.1f}M {100*v/tot:5.1f}%")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
t2 = collections.Counter()
for x in data:
    for h in st:
        if x[ix[h]]:
            t2[h] += float(x[ix[h]])
s = sum(t2.values()) or 1
print("stalls:", {k: round(100 * v / s, 1) for k, v in t2.most_common(8)})
acc, ph = 0, []
for x in data:
    acc += float(x[ix["Instructions Executed"]] or 0)
    if "BAR.SYNC" in x[ix["Source"]]:
        ph.append(acc); acc = 0
ph.append(acc)
print("warp-instructions per barrier-delimited phase (M):", [round(v / 1e6, 1) for v in ph])
