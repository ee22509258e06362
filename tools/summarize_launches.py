"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    agg = collections.OrderedDict()
    for name, ns, grid, block in rows:
        short = re.sub(r"\(.*", "", name)
        d = agg.setdefault(short, [0, 0.0])
        d[0] += 1; d[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# launch list summary: {len(rows)} launches, {tot/1e6:.3f} ms total (per-launch times are cold-cache and serialised: compare shares)\n")
    print("| kernel | launches | total ms | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns/1e6:.3f} | {100*ns/tot:.1f} % | {ns/n/1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
