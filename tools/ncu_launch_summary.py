#!/usr/bin/env python
"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name.

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches.csv \
        --metrics gpu__time_duration.sum python bench.py --ncu-step --no-cpu-baseline
    python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/rNN_ncu_launch_summary.txt

Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's roofline.share_of_step.
"""
import csv
import re
import sys


def main() -> None:
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg: dict = {}
    n = 0
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        ms = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3,
                                              "msecond": 1.0}.get(r[iu], 1e-6)
        name = re.sub(r"\(.*", "", r[ik])[:110]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
        n += 1
    tot = sum(v[0] for v in agg.values())
    print(f"# launches {n} total ms {tot:.1f} (per-launch times are cold-cache/serialised under ncu: compare SHARES)")
    for name, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        print(f"{ms:10.2f} ms {100 * ms / tot:5.1f}%  {c:5d}x  {name}")


if __name__ == "__main__":
    main()
