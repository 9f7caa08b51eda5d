"""Summarise ncu reports for profiles/: a per-kernel launch list (from the gpu__time_duration CSV) and the key counters
of `ncu --set full` captures (.ncu-rep), as markdown.

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rN_launches_summary.md
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep|raw.csv [...] > profiles/rN_ncu_full_summary.md
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    g = defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
        name = re.sub(r"\(.*", "", r[kn])
        name = re.sub(r"^void ", "", name).replace("<unnamed>::", "").replace("vqa::", "")[:70]
        g[name][0] += 1
        g[name][1] += v
        tot += v
    n = sum(v[0] for v in g.values())
    print("Total %.0f us over %d launches (cold-cache, serialised per-launch times: compare shares).\n" % (tot, n))
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(g.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


WANT = [
    ("gpu__time_duration.sum", "time us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor inst %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"),
    ("dram__bytes_read.sum", "DRAM rd MB"),
    ("dram__bytes_write.sum", "DRAM wr MB"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
]


def full(paths):
    for path in paths:
        if path.endswith(".csv"):      # already exported on the GPU box with `ncu -i rep --page raw --csv` (reports are large)
            out = open(path, errors="replace").read()
        else:
            out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, units = rows[0], rows[1]
        cols = [(h.index(m), lbl, m) for m, lbl in WANT if m in h]
        kn = h.index("Kernel Name")
        print("### %s\n" % path.split("/")[-1])
        print("| # | kernel | " + " | ".join(lbl for _, lbl, _ in cols) + " |")
        print("|---|---|" + "---|" * len(cols))
        for i, r in enumerate(rows[2:]):
            name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("<unnamed>::", "").replace("vqa::", "")[:58]
            vals = []
            for ci, lbl, m in cols:
                v = r[ci].replace(",", "")
                try:
                    f = float(v)
                    u = units[ci]
                    if "bytes" in m:
                        f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    if m == "gpu__time_duration.sum":
                        f = f * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                    vals.append("%.1f" % f if f < 1000 else "%.0f" % f)
                except ValueError:
                    vals.append(v)
            print("| %d | `%s` | " % (i, name) + " | ".join(vals) + " |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
