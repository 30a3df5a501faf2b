"""Blackwell-instruction evidence for the built library: per kernel, the count of tensor-core / TMEM / TMA SASS opcodes
(cuobjdump -sass; no GPU needed).

    python tools/sass_table.py > profiles/rNN_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mr_rl_b200", "_lib", "libmr_rl_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "DMMA", "HMMA", "SYNCS",
         "F2FP", "MUFU", "LDGSTS"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][w] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"library: {os.path.relpath(lib, ROOT)}   ({len(per)} sm_100a kernels)")
print("SASS mnemonics: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA),")
print("                DMMA = mma.sync f64 (FP64 tensor pipe), SYNCS = mbarrier ops, LDGSTS = cp.async\n")
print(f"{'kernel':78s} " + " ".join(f"{w:>7s}" for w in WATCH) + "   insts")
for (k, c), name in zip(per.items(), names):
    if not any(c[w] for w in WATCH[:11]):
        continue
    short = re.sub(r"\(.*", "", name).replace("void mr::", "").replace("(anonymous namespace)::", "")[:78]
    print(f"{short:78s} " + " ".join(f"{c[w]:7d}" for w in WATCH) + f" {c['_total']:7d}")
    tot.update(c)
print(f"{'TOTAL (kernels listed)':78s} " + " ".join(f"{tot[w]:7d}" for w in WATCH) + f" {tot['_total']:7d}")
