"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list by kernel
name: launches, time, DRAM bytes read / written (per step and per launch) and the physical DRAM bandwidth of each kernel.

    python tools/traffic_summary.py gpurun_out/step_traffic.csv > profiles/r2_step_traffic.txt
"""
import csv
import re
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}


def main(path, json_out=None):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = defaultdict(dict)  # launch id -> {metric: value}
    names = {}
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1.0)
        except (ValueError, KeyError):
            continue
        per[r["ID"]][r["Metric Name"]] = v
        names[r["ID"]] = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "")
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    tot = [sum(v[j] for v in agg.values()) for j in range(4)]
    print(f"total: {tot[0]} launches, {tot[1] / 1e6:.3f} ms (cold, serialised), DRAM read {tot[2] / 1e9:.3f} GB, written {tot[3] / 1e9:.3f} GB")
    print(f"{'ms':>9} {'%':>5} {'n':>5} {'us/launch':>10} {'rd MB/launch':>13} {'wr MB/launch':>13} {'DRAM GB/s':>10}  kernel")
    for n, (c, ns, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:9.3f} {100 * ns / tot[1]:5.1f} {c:5d} {ns / c / 1e3:10.1f} {rd / c / 1e6:13.2f} {wr / c / 1e6:13.2f} {(rd + wr) / ns:10.1f}  {n[:100]}")
    if json_out:
        import json
        fam = [v for k, v in agg.items() if any(t in k for t in ("gemm_tc_kernel", "wgrad2_kernel", "ffn_bwd_kernel", "ffn_fwd_kernel", "attn_pair_kernel"))]
        out = {
            "source": path, "launches_per_step": tot[0], "step_ms_cold": tot[1] / 1e6,
            "step_dram_read_bytes": tot[2], "step_dram_write_bytes": tot[3],
            "gemm_family_launches": sum(v[0] for v in fam),
            "gemm_family_bytes_per_step": sum(v[2] + v[3] for v in fam),
            "gemm_family_bytes_per_launch": sum(v[2] + v[3] for v in fam) / max(1, sum(v[0] for v in fam)),
            "note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the tcgen05 GEMM launches of ONE eager training step "
                    "(ncu --cache-control all: every launch starts with cold caches, so operands that the real step finds in the 126 MB "
                    "L2 are counted as DRAM reads here -- an upper bound on the step's real traffic)",
            "per_kernel": {k: {"launches": v[0], "ms": v[1] / 1e6, "dram_read_bytes": v[2], "dram_write_bytes": v[3]} for k, v in agg.items() if v[1] / tot[1] > 0.002},
        }
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--json" else None)
