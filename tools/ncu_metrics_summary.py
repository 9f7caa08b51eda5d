"""Summaries of the per-launch ncu metric pass over ONE training step (bench.py --ncu-step):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,lts__t_bytes.sum,...
        --clock-control none --profile-from-start off --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --ncu-step

    python tools/ncu_metrics_summary.py gpurun_out/r2e_launches.csv r2
writes profiles/<tag>_launches_wide.csv (one row per launch), <tag>_launches_summary.md (per kernel), <tag>_conv_summary.md
(every backbone convolution: time, DRAM bytes vs algorithmic bytes, tensor-pipe activity) and <tag>_dram_traffic.json
(bench.py's roofline.traffic).  Per-launch times under ncu are cold-cache and serialised: compare shares and bytes.
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHORT = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct", "lts__t_bytes.sum": "l2_bytes",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
         "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct"}


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        d = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        key = SHORT.get(r[ix["Metric Name"]], r[ix["Metric Name"]])
        if key == "us":
            v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        if unit == "Kbyte":
            v *= 1e3
        elif unit == "Mbyte":
            v *= 1e6
        elif unit == "Gbyte":
            v *= 1e9
        d[key] = v
    return list(launches.values())


def short_name(n):
    n = re.sub(r"\(.*", "", n)
    return re.sub(r"^void ", "", n).replace("<unnamed>::", "").replace("vqa::", "")


def resnet50_conv_table(B=64, H=224):
    """(name, M, Cout, K, algorithmic bytes) of the backbone's convolution launches in plan order (plan_builder.py): stem, then
    per bottleneck conv1, conv2, [downsample], conv3; bf16 activations in / out (+ bf16 residual for conv3), bf16 weights."""
    out = []
    h = H // 2
    out.append(("stem 7x7/2 3->64", B * h * h, 64, 147, B * H * (H + 8) * 8 * 2 + B * h * h * 64 * 2, 2.0 * B * h * h * 64 * 147))
    h //= 2
    cin = 64
    for li, (planes, n) in enumerate(zip([64, 128, 256, 512], [3, 4, 6, 3])):
        for bi in range(n):
            s = 2 if (bi == 0 and li > 0) else 1
            ho = h // s
            pre = "layer%d.%d" % (li + 1, bi)

            def row(nm, hin, hout, ci, co, k, res):
                M = B * hout * hout
                byt = B * hin * hin * ci * 2 + M * co * 2 * (2 if res else 1) + co * ci * k * k * 2
                out.append(("%s %s %dx%d %d->%d k%d%s" % (pre, nm, hout, hout, ci, co, k, " +res" if res else ""), M, co,
                            ci * k * k, byt, 2.0 * M * co * ci * k * k))
            row("conv1", h, h, cin, planes, 1, False)
            row("conv2", h, ho, planes, planes, 3, False)
            if bi == 0:
                row("downsample", h, ho, cin, planes * 4, 1, False)
            row("conv3", ho, ho, planes, planes * 4, 1, True)
            cin, h = planes * 4, ho
    out.append(("projection 7x7 2048->768 k3 (ConvTranspose2d)", B * 49, 768, 2048 * 9,
                B * 49 * 2048 * 2 + B * 49 * 768 * 2 + 768 * 2048 * 9 * 2, 2.0 * B * 49 * 768 * 2048 * 9))
    return out


def simple(path, tag, title):
    """Per-kernel table only (any step: used for the VitVQAModel step, tools/vit_bench.py --ncu-step)."""
    L = load(path)
    tot = sum(d["us"] for d in L)
    g = collections.OrderedDict()
    for d in L:
        e = g.setdefault(short_name(d["name"])[:72], dict(n=0, us=0.0, rd=0.0, wr=0.0, tw=0.0))
        e["n"] += 1; e["us"] += d["us"]; e["rd"] += d.get("dram_rd", 0); e["wr"] += d.get("dram_wr", 0)
        e["tw"] += d.get("tensor_pct", 0) * d["us"]
    with open(os.path.join(ROOT, "profiles", tag + "_launches_summary.md"), "w") as f:
        f.write("# %s\n\nTotal %.0f us over %d launches (ncu: cold-cache, serialised per-launch times: compare shares).\n\n"
                % (title, tot, len(L)))
        f.write("| kernel | launches | total us | share | avg us | DRAM rd MB | DRAM wr MB | GB/s | tensor pipe active % (time-weighted) |\n|---|---|---|---|---|---|---|---|---|\n")
        for k, e in sorted(g.items(), key=lambda kv: -kv[1]["us"]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f | %.1f | %.0f | %.1f |\n" % (
                k, e["n"], e["us"], 100 * e["us"] / tot, e["us"] / e["n"], e["rd"] / 1e6, e["wr"] / 1e6,
                (e["rd"] + e["wr"]) / e["us"] / 1e3, e["tw"] / e["us"]))


def main():
    if len(sys.argv) > 3 and sys.argv[3] == "--simple":
        return simple(sys.argv[1], sys.argv[2], sys.argv[4])
    path, tag = sys.argv[1], sys.argv[2]
    L = load(path)
    prof = os.path.join(ROOT, "profiles")
    keys = ["us", "dram_rd", "dram_wr", "tensor_pct", "l2_bytes", "sm_pct", "dram_pct"]
    with open(os.path.join(prof, tag + "_launches_wide.csv"), "w") as f:
        f.write("idx,kernel,grid," + ",".join(keys) + "\n")
        for i, d in enumerate(L):
            f.write("%d,\"%s\",\"%s\",%s\n" % (i, short_name(d["name"]), d["grid"], ",".join("%.6g" % d.get(k, 0.0) for k in keys)))
    tot = sum(d["us"] for d in L)
    g = collections.OrderedDict()
    for d in L:
        e = g.setdefault(short_name(d["name"])[:72], dict(n=0, us=0.0, rd=0.0, wr=0.0, tw=0.0))
        e["n"] += 1; e["us"] += d["us"]; e["rd"] += d.get("dram_rd", 0); e["wr"] += d.get("dram_wr", 0)
        e["tw"] += d.get("tensor_pct", 0) * d["us"]
    with open(os.path.join(prof, tag + "_launches_summary.md"), "w") as f:
        f.write("# %s: ncu per-launch metrics of ONE training step (R50, batch 64, CUDA graphs on)\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg."
                "pct_of_peak_sustained_active,lts__t_bytes.sum,... --clock-control none --profile-from-start off python bench.py "
                "--steps 2 --warmup 3 --ncu-step` (raw per-launch table: %s_launches_wide.csv).\n\n" % tag)
        f.write("Total %.0f us over %d launches (cold-cache, serialised per-launch times: compare shares; the step itself overlaps "
                "lanes and streams).\n\n" % (tot, len(L)))
        f.write("| kernel | launches | total us | share | avg us | DRAM rd MB | DRAM wr MB | GB/s | tensor pipe active % (time-weighted) |\n|---|---|---|---|---|---|---|---|---|\n")
        for k, e in sorted(g.items(), key=lambda kv: -kv[1]["us"]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f | %.1f | %.0f | %.1f |\n" % (
                k, e["n"], e["us"], 100 * e["us"] / tot, e["us"] / e["n"], e["rd"] / 1e6, e["wr"] / 1e6,
                (e["rd"] + e["wr"]) / e["us"] / 1e3, e["tw"] / e["us"]))
    gem = [d for d in L if "gemm_tcgen05_kernel" in d["name"]]
    table = resnet50_conv_table()
    with open(os.path.join(prof, tag + "_conv_summary.md"), "w") as f:
        f.write("# %s: every convolution launch of the frozen ResNet-50 backbone + the channel projection (batch 64, 224x224)\n\n" % tag)
        f.write("From the same ncu pass as %s_launches_summary.md: the first %d `gemm_tcgen05_kernel` launches of the step are the "
                "implicit-GEMM convolutions in plan order.  `alg MB` = bf16 input + output (+ residual) + weights, the bytes a "
                "perfect kernel moves; `DRAM MB` = dram__bytes_read.sum + dram__bytes_write.sum (activations written by the "
                "previous launch may still be L2-resident: DRAM < alg is possible; DRAM >> alg would mean wasted re-reads).\n\n" % (tag, len(table)))
        f.write("| # | convolution | M x N x K | us | alg MB | DRAM MB | DRAM/alg | DRAM GB/s | alg GB/s | TF/s | tensor pipe % | L2 traffic MB |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        s_us = s_alg = s_dram = s_fl = 0.0
        for i, (row, d) in enumerate(zip(table, gem)):
            name, M, N, K, byt, fl = row
            dram = d.get("dram_rd", 0) + d.get("dram_wr", 0)
            s_us += d["us"]; s_alg += byt; s_dram += dram; s_fl += fl
            f.write("| %d | %s | %d x %d x %d | %.1f | %.1f | %.1f | %.2f | %.0f | %.0f | %.0f | %.1f | %.0f |\n" % (
                i, name, M, N, K, d["us"], byt / 1e6, dram / 1e6, dram / byt, dram / d["us"] / 1e3, byt / d["us"] / 1e3,
                fl / d["us"] / 1e6, d.get("tensor_pct", 0), d.get("l2_bytes", 0) / 1e6))
        f.write("\nSum: %.0f us, algorithmic %.0f MB, DRAM %.0f MB (ratio %.2f), %.0f TF/s over the %d launches; HBM floor of the "
                "algorithmic bytes at 6546 GB/s: %.0f us.\n" % (s_us, s_alg / 1e6, s_dram / 1e6, s_dram / s_alg, s_fl / s_us / 1e6,
                                                               len(table), s_alg / 6546e3))
    fam_bytes = sum(d.get("dram_rd", 0) + d.get("dram_wr", 0) for d in gem)
    adam = [d for d in L if "adamw_kernel" in d["name"]]
    adam_bytes = sum(d.get("dram_rd", 0) + d.get("dram_wr", 0) for d in adam)
    with open(os.path.join(prof, tag + "_dram_traffic.json"), "w") as f:
        json.dump({"source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one training step, profiles/%s_launches_wide.csv" % tag,
                   "gemm_family_bytes_per_step": fam_bytes, "gemm_family_launches": len(gem),
                   "adamw_bytes_per_step": adam_bytes, "adamw_launches": len(adam),
                   "adamw_bytes_per_launch": adam_bytes,   # bench.py times ONE launch over all parameters: same bytes
                   "step_dram_bytes": sum(d.get("dram_rd", 0) + d.get("dram_wr", 0) for d in L)}, f, indent=1)
    print("launches", len(L), "gemm launches", len(gem), "total us %.0f" % tot)


if __name__ == "__main__":
    main()
