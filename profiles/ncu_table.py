"""One line per distinct kernel of an ncu --set full report: time, DRAM bytes, issue/occupancy, registers, fp64 pipe, cache hit rates.
usage: python profiles/ncu_table.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum"]
ix = [hdr.index(k) for k in keys]
def f(x):
    try: return float(x.replace(",", ""))
    except ValueError: return float("nan")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
bscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
print(f"{'kernel':24s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'issue%':>6s} {'warps%':>6s} {'regs':>4s} {'fp64%':>5s} {'L1hit':>5s} {'L2hit':>5s} {'Minst':>7s}")
seen = set()
for r in rows[2:]:
    n = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    if n in seen: continue
    seen.add(n)
    v = [r[i] for i in ix]
    t = f(v[0]) * scale.get(units[ix[0]], 1.0)
    rd = f(v[1]) * bscale.get(units[ix[1]], 1.0); wr = f(v[2]) * bscale.get(units[ix[2]], 1.0)
    print(f"{n:24s} {t:8.1f} {rd:8.1f} {wr:8.1f} {(rd+wr)/t*1e3:7.0f} {f(v[3]):6.1f} {f(v[4]):6.1f} {v[5]:>4s} {f(v[6]):5.1f} {f(v[7]):5.1f} {f(v[8]):5.1f} {f(v[9])/1e6:7.1f}")
