"""DRAM traffic of ONE d_sw call from an ncu metrics list
(ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv ... python profiles/prof_dsw.py).
Sums the kernels of the LAST complete d_sw call (from one k_dsw_wind launch to the next).
usage: dsw_traffic.py list.csv [--json profiles/r2_dsw_traffic.json KEY]   (KEY e.g. C384L79A: the entry bench.py reads for roofline.traffic)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mn, mu, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    d = launch.setdefault(int(r[idc]), {"name": r[kn].split("(")[0].replace("void ", "")})
    d[r[mn]] = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
ids = list(launch)
winds = [i for i in ids if launch[i]["name"].startswith("k_dsw_wind")]
lo, hi_ = winds[-2], winds[-1]
sel = [launch[i] for i in ids if lo <= i < hi_]
rd = sum(x.get("dram__bytes_read.sum", 0) for x in sel); wr = sum(x.get("dram__bytes_write.sum", 0) for x in sel)
t = sum(x.get("gpu__time_duration.sum", 0) for x in sel)
print(f"one d_sw call: {len(sel)} launches, {t/1e3:.3f} ms (serialised), DRAM read {rd/1e9:.3f} GB + write {wr/1e9:.3f} GB = {(rd+wr)/1e9:.3f} GB")
agg = collections.OrderedDict()
for x in sel:
    a = agg.setdefault(x["name"], [0, 0., 0., 0.]); a[0] += 1; a[1] += x.get("gpu__time_duration.sum", 0); a[2] += x.get("dram__bytes_read.sum", 0); a[3] += x.get("dram__bytes_write.sum", 0)
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:32s} x{a[0]} {a[1]:8.1f} us  rd {a[2]/1e6:8.1f} MB  wr {a[3]/1e6:8.1f} MB")

if "--json" in sys.argv:
    import json, os, subprocess
    out, key = sys.argv[sys.argv.index("--json") + 1], sys.argv[sys.argv.index("--json") + 2]
    rec = json.load(open(out)) if os.path.exists(out) else {}
    rec[key] = {"bytes": rd + wr, "read": rd, "write": wr, "launches": len(sel), "ncu_serialised_ms": t / 1e3, "csv": os.path.basename(sys.argv[1]),
                "kernels": {n: {"launches": a[0], "us": round(a[1], 1), "read_MB": round(a[2] / 1e6, 1), "write_MB": round(a[3] / 1e6, 1)} for n, a in agg.items()}}
    json.dump(rec, open(out, "w"), indent=1)
