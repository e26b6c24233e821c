"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (last N launches = one step).

    python tools/launch_summary.py gpurun_out/launches.csv [n_last] > profiles/rN_..._launches.txt
"""
import collections
import csv
import re
import sys


def main(path, n_last=None):
    rows = []
    with open(path) as f:
        for line in f:
            if line.startswith('"ID"'):
                break
        for r in csv.reader(f):
            if len(r) >= 15 and r[12] == "gpu__time_duration.sum":
                name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
                rows.append((name, float(r[14]) / 1000.0))
    if n_last:
        rows = rows[-int(n_last):]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1000:.3f} ms over {len(rows)} launches")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1000:9.3f} ms {100 * us / tot:5.1f}% {c:5d}x {us / c:9.1f} us  {n[:110]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
