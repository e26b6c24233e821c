"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (bulk tensor load / store / reduce), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus HMMA / IMMA (legacy mma.sync: expected 0).

    python tools/sass_summary.py [liteasr_b200/liblasr.so] > profiles/r2_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "liteasr_b200/liblasr.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "IMMA", "REDG", "ATOMG", "MUFU"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"SASS summary of {lib} (sm_100a): {len(per)} kernels")
print(f"{'instr':>7} " + " ".join(f"{k:>8}" for k in KEYS) + "  kernel")
rows = []
for (name, c), d in zip(per.items(), dem):
    tot.update(c)
    short = re.sub(r"\(.*", "", d).replace("void ", "")
    rows.append((c, short))
for c, short in sorted(rows, key=lambda r: -(r[0]["UTCHMMA"] * 1000 + r[0]["UTMALDG"])):
    if c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"]:
        print(f"{c['_total']:7d} " + " ".join(f"{c[k]:8d}" for k in KEYS) + f"  {short[:90]}")
print(f"{tot['_total']:7d} " + " ".join(f"{tot[k]:8d}" for k in KEYS) + "  TOTAL (all kernels, incl. the bandwidth kernels not listed above)")
