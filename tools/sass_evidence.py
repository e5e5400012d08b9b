"""SASS evidence of the Blackwell path: per kernel of the built library, the count of every tensor-core / TMEM / TMA /
mbarrier / multimem mnemonic and its first occurrence with operands.

  python tools/sass_evidence.py > profiles/r2_sass.txt         (needs cuobjdump and c++filt; no GPU)

tcgen05.mma -> UTCHMMA (kind::tf32 and kind::f16 share the mnemonic; .2CTA = cta_group::2), tcgen05.ld/st -> LDTM/STTM,
tcgen05.commit -> UTCBAR (.2CTA.MULTICAST for the pair), tcgen05.alloc -> UTCATOMSWS, cp.async.bulk.tensor -> UTMALDG,
mbarrier -> SYNCS.*, cvt.rn.bf16x2.f32 -> F2FP.BF16.F32.PACK_AB, multimem.ld_reduce -> LDGMC.*.ADD (multimem.st is an
ordinary STG on the multicast address), system-scope CAS of the rank barrier -> ATOMG.E.CAS.STRONG.SYS."""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorchhessianfree_b200", "csrc", "libhf_b200.so")
WANT = re.compile(r"^(UTC|LDTM|STTM|UTMA|SYNCS|F2FP\.BF16|LDGMC|ATOMG\.E\.CAS\.STRONG\.SYS|UBLKCP|MEMBAR\.ALL\.SYS)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?;)", line)
        if m and cur is not None:
            cur.append(m.group(1).strip())
    names = list(funcs)
    pretty = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print(__doc__.strip().replace("\n", "\n# ").join(["# ", ""]))
    for mangled, name in zip(names, pretty):
        counts, first = OrderedDict(), OrderedDict()
        for ins in funcs[mangled]:
            body = re.sub(r"^@!?U?P\d+\s+", "", ins)
            op = body.split()[0]
            if WANT.match(op):
                counts[op] = counts.get(op, 0) + 1
                first.setdefault(op, ins)
        if not counts:
            continue
        print(f"\n== {name}")
        print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(counts.items())))
        for ins in first.values():
            print("      " + ins)


if __name__ == "__main__":
    sys.exit(main())
