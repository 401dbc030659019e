"""ncu launch list (gpu__time_duration.sum) of a whole command -> per-kernel launches / time / share (markdown).
   usage: python profiles/launch_shares.py list.csv "description" > shares.md"""
import collections, csv, re, sys
rows = list(csv.DictReader([l for l in open(sys.argv[1], errors="ignore") if not l.startswith("==")]))
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values()); nl = sum(v[0] for v in agg.values())
print(f"ncu launch list of {sys.argv[2]}: {nl} launches, {tot / 1e6:.1f} ms of kernel time (cold-cache, serialised: compare SHARES)\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t / 1e6:.2f} | {100 * t / tot:.1f} % |")
stage = collections.OrderedDict([("d_sw", ("k_dsw", "k_deln", "k_copy_frame")), ("c_sw", ("k_csw",)), ("column solvers", ("k_riem",)),
                                 ("update_dz", ("k_tp_zn", "k_tp_fused", "k_edge", "k_dz", "k_dzc")), ("pressure gradient", ("k_a2b", "k_nh_pgrad", "k_pgrad_c", "k_pk3", "k_pe_halo")),
                                 ("halo exchange", ("k_halo", "k_p2p"))])
print("\n| stage (kernel-name prefix) | share |\n|---|---|")
for s, pre in stage.items():
    t = sum(v[1] for k, v in agg.items() if k.startswith(pre))
    print(f"| {s} | {100 * t / tot:.1f} % |")
