"""SASS mnemonic census of the shipped library (no GPU needed): proves which kernels use tcgen05 / TMA / TMEM / clusters.
    python tests/sass_census.py > profiles/r2_sass_census.txt"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "parseq_b200", "lib", "libparseq_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR",
      "HMMA", "LDSM", "FFMA2", "FADD2", "MUFU.EX2", "LDGSTS", "ST.E", "STS", "MAPA", "CCTL", "ACQBULK", "BAR.SYNC", "ERRBAR", "MEMBAR"]
cur, per = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*$", "", name).replace("void pq::", "").replace("pq::", "")
        cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for k in MN:
            if op.startswith(k):
                cur[k] += 1
print("# SASS mnemonic census of parseq_b200/lib/libparseq_b200.so (sm_100a), cuobjdump -sass; per kernel instantiation")
print("# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, LDTM / STTM = tcgen05.ld / st (TMEM),")
print("# SYNCS = mbarrier ops, UCGABAR = barrier.cluster, MAPA = DSMEM address mapping, HMMA / LDSM = mma.sync / ldmatrix, FFMA2 = packed fp32x2 FMA")
cols = [k for k in MN if any(c[k] for c in per.values())]
print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{k[:8]:>8s}" for k in cols))
tot = collections.Counter()
for name, c in per.items():
    print(f"{name[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[k]:8d}" for k in cols))
    tot.update(c)
print(f"{'TOTAL':58s} {tot['_total']:7d} " + " ".join(f"{tot[k]:8d}" for k in cols))
