"""Per-function SASS comparison of two builds of the library (cuobjdump -sass): which kernels are byte-identical, which differ,
which are new.  Used to prove that a refactor (host annotations, extra template instantiations, new schemes in a separate
instantiation) leaves the GPU-validated kernels untouched when no GPU time is available to re-measure them.
   usage: python profiles/sasscmp.py validated/libfv3_b200.so gfdl_atmos_cubed_sphere_b200/csrc/libfv3_b200.so [old_name=new_name ...]
   (old_name=new_name compares a renamed kernel, e.g. a function that became a template instantiation; substrings of the mangled names)"""
import collections
import difflib
import re
import subprocess
import sys


def dump(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, d = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1); d[fn] = []; continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            d[fn].append(re.sub(r"/\*[0-9a-f]{4,6}\*/", "", line).strip())
    return d


def main():
    a, b = dump(sys.argv[1]), dump(sys.argv[2])
    print(len(a), "functions in", sys.argv[1], "/", len(b), "in", sys.argv[2])
    same = True
    for fn, ins in a.items():
        if fn not in b:
            print("  missing in new :", fn[:110])
        elif b[fn] != ins:
            same = False
            print("  DIFFERS        :", fn[:110], len(ins), "->", len(b[fn]))
    for fn in b:
        if fn not in a:
            print("  new function   :", fn[:110], len(b[fn]), "instructions")
    for pair in sys.argv[3:]:
        o, n = pair.split("=")
        fo = [f for f in a if o in f]; fn_ = [f for f in b if n in f]
        if len(fo) == 1 and len(fn_) == 1:
            eq = a[fo[0]] == b[fn_[0]]
            print(f"  renamed {o} -> {n}: {'IDENTICAL' if eq else 'DIFFERENT'} ({len(a[fo[0]])} / {len(b[fn_[0]])} instructions)")
            if not eq:
                norm = lambda l: re.sub(r"0x[0-9a-f]+", "X", re.sub(r"/\*.*?\*/", "", l)).strip()
                for l in list(difflib.unified_diff([norm(x) for x in a[fo[0]]], [norm(x) for x in b[fn_[0]]], lineterm="", n=1))[:40]:
                    print("     ", l)
        else:
            print(f"  renamed {o} -> {n}: ambiguous or not found ({len(fo)} / {len(fn_)} matches)")
    print("ALL COMMON FUNCTIONS IDENTICAL" if same else "SOME COMMON FUNCTIONS DIFFER")


if __name__ == "__main__":
    main()
