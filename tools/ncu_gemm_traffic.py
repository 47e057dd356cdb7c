#!/usr/bin/env python
"""Summarise a strided ncu sample of the tensor-core GEMM launches of one training step.

Produced on the GPU box by
    NK_BENCH_MIN_WARMUP=0 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/gemm_traffic.csv \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        python bench.py --ncu-sample 8 --no-cpu-baseline
(bench.py puts every 8th gemm_tc launch of one eager step inside a cudaProfilerStart/Stop range and writes the sampled
launches' shapes to gpurun_out/gemm_traffic_shapes.json).

usage: python tools/ncu_gemm_traffic.py gemm_traffic.csv gemm_traffic_shapes.json out.json
"""
import csv
import json
import sys


def algorithmic_bytes(what: str, ints: list):
    if what.startswith("linear") and len(ints) >= 3:
        m, n, k = ints[-3:]
        if what == "linear_fwd":
            return 2 * (m * k + n * k) + 2 * m * n
        if what == "linear_dgrad":
            return 2 * (m * n + n * k) + 2 * m * k
        if what == "linear_wgrad":
            return 2 * (m * n + m * k) + 4 * n * k
    if what.startswith("conv2d") and len(ints) >= 6:
        nimg, h, w, ci, co, ks = ints[-6:]
        px = nimg * h * w
        if what == "conv2d_fwd":
            return 2 * px * ci + 2 * co * ci * ks * ks + 2 * px * co
        if what == "conv2d_wgrad":
            return 2 * px * (ci + co) + 4 * co * ci * ks * ks
    return None


def main() -> None:
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    iu = hdr.index("Metric Unit")
    per: dict = {}
    for r in rows[1:]:
        if "gemm_tc_kernel" not in r[ik]:
            continue
        val = float(r[iv].replace(",", ""))
        unit = r[iu]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
                 "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)
        per.setdefault(int(r[iid]), {})[r[im]] = val * scale
    launches = [per[k] for k in sorted(per)]
    shapes = json.loads(open(sys.argv[2]).read())
    dram = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in launches]
    dur = [d.get("gpu__time_duration.sum", 0.0) for d in launches]
    alg, flops = [], []
    for what, fl, ints in shapes["launches"]:
        alg.append(algorithmic_bytes(what, ints))
        flops.append(fl)
    n = len(launches)
    pairs = [(a, b) for a, b in zip(alg, dram) if a is not None] if len(alg) == n else []
    out = {
        "source": "ncu --profile-from-start off, every %d-th gemm_tc launch of one eager step, batch %d per GPU"
                  % (shapes["every"], shapes["batch_per_gpu"]),
        "sampled_launches": n,
        "dram_bytes_per_launch": sum(dram) / max(1, n),
        "duration_us_per_launch_under_ncu": sum(dur) / max(1, n),
        "algorithmic_bytes_per_launch": (sum(a for a, _ in pairs) / len(pairs)) if pairs else None,
        "dram_bytes_per_launch_same_subset": (sum(b for _, b in pairs) / len(pairs)) if pairs else None,
        "tflop_per_launch": (sum(flops) / len(flops) / 1e12) if flops else None,
    }
    if pairs:
        out["traffic_over_algorithmic"] = out["dram_bytes_per_launch_same_subset"] / out["algorithmic_bytes_per_launch"]
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
