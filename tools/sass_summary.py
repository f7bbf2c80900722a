#!/usr/bin/env python
"""Counts of the Blackwell-specific SASS mnemonics per kernel of libhode.so (cuobjdump -sass): the evidence that the hot
kernels are tcgen05 / TMEM / bulk-copy code.  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hybrid_ode_for_glp_1_and_glucose_b200", "libhode.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCATOMSWS", "SYNCS", "LDGSTS", "HMMA", "FFMA", "MUFU", "ELECT", "USETMAXREG"]
cur, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"arch = (\S+)", line.strip())
    if m:
        arch = m.group(1)
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            for k in MN:
                if op == k:
                    counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (arch {arch}); counts of selected SASS mnemonics per kernel")
print(f"# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine),")
print(f"# SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, UTMALDG = tensor-map TMA (not used: operands are linear images)")
hdr = ["kernel", "instr"] + MN
print("  ".join(f"{h:>10s}" if i else f"{h:70s}" for i, h in enumerate(hdr)))
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
for (k, c), name in zip(counts.items(), dem or counts):
    short = (name.split(">(")[0] + ">" if ">(" in name else re.sub(r"\(.*", "", name)).replace("(int)", "")[:70]
    print(f"{short:70s}  {c['_total']:>10d}  " + "  ".join(f"{c[m]:>10d}" for m in MN))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL':70s}  {tot['_total']:>10d}  " + "  ".join(f"{tot[m]:>10d}" for m in MN))
