"""Condense `cuobjdump -sass libd3p_b200.so` into profiles/sass_mnemonics.txt: per kernel the SASS instruction count and the
counts of the tensor-core, TMA, mbarrier, warp-MMA, memory and reduction mnemonics (no GPU needed).

    python scripts/sass_mnemonics.py > profiles/sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "d3p_b200", "_lib", "libd3p_b200.so")
PAT = re.compile(r"\b(UTCHMMA[.\w]*|UTMALDG[.\w]*|UTMACCTL[.\w]*|LDTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|HMMA[.\w]*"
                 r"|ELECT\b|PREEXIT|ACQBULK|REDUX[.\w]*|MUFU[.\w]*|LDG[.\w]*|STG[.\w]*|RED[.\w]*|ATOMG[.\w]*|STS[.\w]*|LDS[.\w]*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, per, n = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            n[cur] += 1
            mm = PAT.search(line)
            if mm:
                per[cur][mm.group(1)] += 1
    names = list(per)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass d3p_b200/_lib/libd3p_b200.so, condensed: per kernel the SASS instruction count and the counts of the")
    print("# tensor-core (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc),")
    print("# TMA (UTMALDG = cp.async.bulk.tensor, UTMACCTL = prefetch.tensormap), mbarrier (SYNCS), warp-MMA (HMMA), memory and")
    print("# reduction mnemonics.  Regenerate: python scripts/sass_mnemonics.py > profiles/sass_mnemonics.txt (no GPU needed).")
    for nm, d in zip(names, dem):
        if not n[nm]:
            continue
        short = re.sub(r"\(.*", "", d).replace("d3p::", "").replace("(anonymous namespace)::", "")
        c = per[nm]
        keys = sorted(c, key=lambda k: (-c[k], k))
        print(f"{short[:140]}  [{n[nm]} instr]\n    " + "  ".join(f"{k}:{c[k]}" for k in keys))


if __name__ == "__main__":
    main()
