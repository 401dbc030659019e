"""Per-source-line warp-instruction counts of one kernel from an ncu report (needs -lineinfo and --import-source on).
usage: python profiles/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]   # rx: kernel-name regex, or id:<n> for the n-th profiled launch
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", *(["--kernel-id", ":::" + rx[3:]] if rx.startswith("id:") else ["--kernel-name", "regex:" + rx])],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, seen_fn, hdr = None, None, None
agg = collections.OrderedDict()
first_fn = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if first_fn is None: first_fn = r[1]
        seen_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if seen_fn != first_fn or hdr is None or not r[0].isdigit(): continue
    ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
    try: n = float(r[ie]); s = float(r[sm] or 0)
    except ValueError: continue
    key = (fname, int(r[0]))
    if key not in agg: agg[key] = [0., 0., r[1].strip()]
    agg[key][0] += n; agg[key][1] += s
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values()) or 1
print(f"{first_fn}: {tot/1e6:.1f}M warp-instructions")
byfile = collections.Counter()
for (f, l), v in agg.items(): byfile[f] += v[0]
print("by file:", {k: f"{100*v/tot:.1f}%" for k, v in byfile.most_common()})
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.1f}% inst {100*v[1]/ts:5.1f}% smp  {f}:{l:<4d} {v[2][:110]}")
