#!/usr/bin/env python
"""One table of every roofline fraction in the committed bench lines (kernel, bound, achieved, peak and where the peak comes from).

    python tools/roofline_table.py > profiles/r2_roofline_table.md
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    return json.load(open(os.path.join(P, name)))


def row(src, workload, value, rf, extra=""):
    peak_src = rf.get("peak_source", "MEASURED_PEAKS.json" if rf.get("bound") in ("hbm", "tensor") else "")
    return (f"| `{src}` | {workload} | {value / 1e6:.1f} M | {rf.get('kernel', '')} | {rf.get('bound', 'hbm')} | "
            f"{rf['achieved']:.0f} {rf.get('unit', 'GB/s')} | {rf.get('peak', float('nan')):.0f} | **{rf['frac']:.2f}** | {peak_src} {extra}|")


def main():
    print("# Roofline fractions of the committed round-2 bench lines\n")
    print("`achieved` = algorithmic bytes (or flops) per launch ÷ the launch's CUDA-event time inside the bench loop; `peak` = the measured "
          "number of `MEASURED_PEAKS.json` (HBM copy 6572.5 GB/s, cuBLAS bf16 1455 TFLOP/s sustained) unless the last column says otherwise.  "
          "Produced by `tools/roofline_table.py` from the JSON lines named in the first column.\n")
    print("| line | workload | samples/s | kernel | bound | achieved | peak | frac | peak source / note |")
    print("|---|---|---|---|---|---|---|---|---|")
    d = load("r2_bench_default.json")
    print(row("r2_bench_default.json", "config 5, 1 GPU (headline)", d["value"], d["roofline"]))
    rb = dict(d["roofline_bwd"], bound="hbm", unit="GB/s", peak=d["roofline"]["peak"])
    print(row("r2_bench_default.json", "config 5, 1 GPU", d["value"], rb))
    rs = dict(d["roofline_step"], bound="hbm", unit="GB/s", peak=d["roofline"]["peak"], kernel="whole step (fwd + bwd)")
    print(row("r2_bench_default.json", "config 5, 1 GPU", d["value"], rs, f"{d['roofline_step']['bytes_per_sample']} B/sample "))
    z = d.get("zipf")
    if z:
        print(f"| `r2_bench_default.json` | config 5, Zipf(1.05) ids | {z['value'] / 1e6:.1f} M | embed_fm2_fwd_kernel | L2-served | "
              f"{z['fwd_algorithmic_GBps']:.0f} GB/s (algorithmic) | {d['roofline']['peak']:.0f} | {z['fwd_frac_of_hbm_peak']:.2f} | not an HBM utilisation: hot rows never reach HBM |")
    for name, c in d["configs"].items():
        rf = c["roofline"]
        extra = ""
        if "frac_of_3xtf32_ceiling" in rf:
            extra = f"; {rf['frac_of_3xtf32_ceiling']:.2f} of the 3xTF32 ceiling (peak/6) "
        if "step_frac" in rf:
            extra = f"; whole step {rf['step_frac']:.2f} (launch-bound: 79 MB per step) "
        print(row("r2_bench_default.json", name, c["value"], rf, extra))
    for n in (2, 4, 8):
        f = f"r2_bench_sharded_n{n}_final.json"
        if not os.path.exists(os.path.join(P, f)):
            continue
        s = load(f)
        rf = s["roofline"]
        print(row(f, f"config 5 row-sharded, {n} GPUs", s["value"], rf))
        push = dict(rf["push"], bound="nvlink", unit="GB/s", peak=rf["peak"], peak_source=rf.get("peak_source", ""))
        print(row(f, f"config 5 row-sharded, {n} GPUs", s["value"], push))


if __name__ == "__main__":
    main()
