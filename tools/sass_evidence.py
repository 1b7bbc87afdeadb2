"""Instruction-mnemonic counts per kernel from `cuobjdump -sass` of the built library (runs without a GPU).
Usage: python tools/sass_evidence.py > profiles/rNN/sass_evidence.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = next(ROOT.glob("*_b200/libsdpa_b200.so"))
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "UCGABAR", "MUFU.EX2", "FFMA2", "FADD2",
         "F2FP", "F2F", "ELECT", "STG.E.128", "LDG.E.128", "SHFL", "USETMAXREG", "NANOSLEEP"]
print("SASS evidence (cuobjdump -sass libsdpa_b200.so, sm_100a) -- instruction counts per kernel")
print("mnemonics: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk (1-D bulk copy),")
print("LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier, UCGABAR = cluster barrier, MUFU.EX2 = ex2.approx,")
print("FFMA2/FADD2 = packed f32x2, USETMAXREG = setmaxnreg\n")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for blk, name in zip(blocks, names):
    short = re.sub(r"\(anonymous namespace\)::|sdpa::", "", name.split("(")[0] if "<" not in name else name[:name.rfind(">") + 1])
    ops = collections.Counter()
    total = 0
    for line in blk.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        total += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                ops[w] += 1
                break
    print(short)
    print(f"    instructions {total}: " + ", ".join(f"{w} {ops[w]}" for w in WATCH if ops[w]))
