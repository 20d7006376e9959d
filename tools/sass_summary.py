"""Per-kernel SASS census of the built library (CPU only: cuobjdump).  Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "r2dm_b200", "libr2dm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
pats = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
        ("LDG/STG.256", r"\b(LDG|STG)\.E\.(ENL2\.)?256"), ("RED/ATOM", r"\b(RED|ATOM|REDG|ATOMG|ATOMS)\b"),
        ("MUFU.TANH", r"MUFU\.TANH"), ("MUFU.EX2", r"MUFU\.EX2"), ("F*2 (packed fp32)", r"\b(FFMA2|FADD2|FMUL2)"),
        ("HMMA", r"\bHMMA")]
print("# SASS census of r2dm_b200/libr2dm_b200.so (cuobjdump -sass, sm_100a).  UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,")
print("# UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, F*2 = packed fp32x2 arithmetic, HMMA = legacy mma.sync (none expected).")
print("kernel | instructions | " + " | ".join(n for n, _ in pats))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    body = blk.split("\n", 1)[1] if "\n" in blk else ""
    ins = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    if "r2dm" not in name:
        continue
    txt = "\n".join(ins)
    print(f"{name[:110]} | {len(ins)} | " + " | ".join(str(len(re.findall(p, txt))) for _, p in pats))
