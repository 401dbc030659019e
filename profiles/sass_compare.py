"""Compare two `cuobjdump -sass` dumps kernel by kernel (instruction text and encodings; anonymous-namespace hashes in the mangled
names are normalised).  Used at the end of round 2, when the GPU budget was spent, to show that adding the ppm_profile
instantiations and k_fillz left every kernel of the B200-validated library unchanged:

    cuobjdump -sass <validated>/libfv3_b200.so > a.sass; cuobjdump -sass <current>/libfv3_b200.so > b.sass
    python profiles/sass_compare.py a.sass b.sass
    -> 183 before, 184 after; changed: none ; new: ['..k_fillz..']
"""
import re
import sys


def funcs(path):
    out, cur = {}, None
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_(\d+)_(\w+?)_cu_[0-9a-f]{8}", r"ANON_\2", m.group(1))
            out[cur] = []
            continue
        if cur and line.strip().startswith("/*"):
            out[cur].append(line.strip())
    return out


if __name__ == "__main__":
    b, a = funcs(sys.argv[1]), funcs(sys.argv[2])
    bad = [n for n in b if a.get(n) != b[n]]
    print(len(b), "before,", len(a), "after; changed:", bad if bad else "none", "; new:", [n for n in a if n not in b])
    sys.exit(1 if bad else 0)
