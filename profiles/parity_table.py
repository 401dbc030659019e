"""FV3_PARITY_LOG record (tests/harness.py::compare) -> markdown table of the worst field per test.
   usage: python profiles/parity_table.py log.jsonl > profiles/r2_parity_errors.md"""
import collections, json, sys
rec = collections.OrderedDict()
n = 0
for line in open(sys.argv[1]):
    d = json.loads(line); n += 1
    t = d["test"].split("::")[-1].split(" ")[0]
    e = rec.setdefault(t, {})
    for k, v in d["err"].items():
        e[k] = max(e.get(k, 0.0), v)
print("# Round 2 — per-field parity errors of one full `pytest -m gpu` run on a B200 (CUDA vs CPU oracle)\n")
print(f"Recorded with `FV3_PARITY_LOG=… python -m pytest tests -m gpu` (`tests/harness.py::compare` appends every comparison; {n} comparisons);")
print("table by `profiles/parity_table.py`.  Error = max |cuda − oracle| / max |oracle| over the compared region.  The tolerances in the")
print("tests (1e-12 per call, 1e-10 per run, documented exceptions for w / ws / long runs) were set from this table.\n")
print("| test | worst field | max error | fields above 1e-12 |\n|---|---|---|---|")
for t, e in rec.items():
    k = max(e, key=e.get)
    above = ", ".join(f"{a} {b:.1e}" for a, b in e.items() if b > 1e-12) or "—"
    print(f"| `{t}` | {k} | {e[k]:.1e} | {above} |")
