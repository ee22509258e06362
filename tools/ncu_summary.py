"""Summarise an ncu --set full report (.ncu-rep) as a markdown table of the metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/conv.ncu-rep > profiles/rNN_ncu_conv_summary.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__cluster_dim_z", "cluster z"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("smsp__cycles_active.avg", "SMSP cycles active (avg)"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{path}` ({len(data)} launches)\n")
    print("| # | kernel | grid | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|---|" + "---|" * len(WANT))
    for k, r in enumerate(data):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for key, _ in WANT:
            if key in idx:
                v, u = r[idx[key]], units[idx[key]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("n/a")
        print(f"| {k} | `{name}` | {r[idx['Grid Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
