"""Generate tests/golden/ppm_notebook.npz from the reference's own NumPy restatement of the
interior xppm (docs/examples/tp_core.ipynb, the "Integration loop" cell).

Run in the build container only (needs /root/reference):  python tests/golden/make_ppm_golden.py
The notebook's numerical block (from '#begin xppm' up to 'flux = flux*c') is exec'ed verbatim,
one step per case, for ord in {5, 6, 8, 10} (+ the positive-definite variant of ord 5 = hord -5),
Courant numbers of both signs.  Cases where the notebook deliberately differs from tp_core.F90
(its smt5 uses '<=' "for graphical purpose") are avoided by using inputs without exact ties.
"""
import json
import os
import textwrap

import numpy as np

NB = "/root/reference/docs/examples/tp_core.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ppm_notebook.npz")


def notebook_blocks():
    nb = json.load(open(NB))
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    consts = next(c for c in cells if c.startswith("#Constants"))
    idx = next(c for c in cells if c.startswith("#Define indices"))
    loop = next(c for c in cells if "#begin xppm" in c)
    body = loop[loop.index("#begin xppm"):loop.index("flux = flux*c")]
    return consts, idx, textwrap.dedent("    " + body)


def main():
    consts, idx, body = notebook_blocks()
    rng = np.random.default_rng(20241117)
    nx = 48
    x = (np.arange(nx) + 0.5) / nx
    inputs = {
        "gauss": np.exp(-((x - 0.5) / 0.12) ** 2) + 0.01 * rng.standard_normal(nx),
        "rough": 1.0 + 0.5 * np.sin(2 * np.pi * x) + 0.2 * rng.standard_normal(nx),
        "steps": np.where((x > 0.3) & (x < 0.6), 1.0, 0.1) + 0.05 * rng.standard_normal(nx),
    }
    out = {}
    for name, q0 in inputs.items():
        for (ordv, PD) in ((5, False), (5, True), (6, False), (8, False), (10, False)):
            for ctag, cval in (("pos", 0.37), ("neg", -0.42), ("mix", None)):
                env = {"np": np, "nx": nx, "ord": ordv, "PD": PD, "lim_fac": 1.0}
                exec(consts, env)
                exec(idx, env)
                q = q0.copy()
                if PD:
                    q = np.abs(q)
                c = np.full(nx + 1, cval) if cval is not None else rng.uniform(-0.6, 0.6, nx + 1)
                env.update(q=q.copy(), c=c.copy())
                exec(body, env)
                key = f"{name}_ord{ordv}{'pd' if PD else ''}_{ctag}"
                out[key + "_q"] = q
                out[key + "_c"] = c
                out[key + "_flux"] = np.asarray(env["flux"], dtype=np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out) // 3, "cases")


if __name__ == "__main__":
    main()
