"""Developer tool: top stall-sample SASS lines of an `ncu --page source --csv --print-source sass` dump."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for b in blocks[:1] if kern is None else [x for x in blocks if kern in x["name"]][:1]:
    h = b["hdr"]
    idx = {n: i for i, n in enumerate(h)}
    samp = idx["# Samples"]
    tot = sum(int(r[samp] or 0) for r in b["rows"])
    print(b["name"][:100], "total samples", tot, "sass lines", len(b["rows"]))
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    agg = {h[i]: sum(int(r[i] or 0) for r in b["rows"]) for i in stall_cols}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    ranked = sorted(enumerate(b["rows"]), key=lambda kv: -int(kv[1][samp] or 0))[:top]
    for pos, r in sorted(ranked):
        st = {h[i][6:]: int(r[i]) for i in stall_cols if r[i] and int(r[i])}
        print(f"{pos:5d} {int(r[samp]):6d} {100 * int(r[samp]) / max(tot, 1):5.1f}%  {r[idx['Source']][:90]:90s} {st}")
