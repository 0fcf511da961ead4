"""Histogram of the Blackwell-specific SASS opcodes per kernel of libtokenhmr_b200.so (cuobjdump -sass):
tcgen05 MMA (UTCHMMA / .2CTA), TMEM loads / stores (LDTM / STTM), TMA (UTMALDG / UTMASTG / UTMAREDG / UTMAPF),
packed fp32 (FFMA2 / FMUL2 / FADD2).  usage: python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "tokenhmr_b200" / "libtokenhmr_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTMACCTL",
        "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "REDG", "ATOMG"]
total = collections.Counter()
print(f"# {lib.name}: SASS opcode counts per kernel (sm_100a), {len(txt.splitlines())} lines of disassembly")
print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA "
      "load / store / reduce (cp.async.bulk.tensor, cp.reduce.async.bulk.tensor), FFMA2 = fma.rn.f32x2\n")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
    c = collections.Counter()
    for o in ops:
        for k in KEYS:
            if o == k or o.startswith(k + "."):
                if k == "UTCHMMA" and ".2CTA" in o:
                    k = "UTCHMMA.2CTA"
                c[k] += 1
                break
    if not c:
        continue
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(.*", "", dem).replace("void ", "")
    print(f"{dem[:100]:100s} insts {len(ops):6d} | " + "  ".join(f"{k} {v}" for k, v in c.items()))
    total.update(c)
print("\nTOTAL " + "  ".join(f"{k} {v}" for k, v in total.most_common()))
