#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for ONE step of bench.py.

    python tools/launch_table.py gpurun_out/launches.csv [launches_per_step]

The step is taken from the END of the list (the last complete `infer`), so warm-up, packing and profiling passes do not
count.  Durations under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"svb::<unnamed>::|svb::|void |at::native::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:86]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    launches = []
    for r in rows[hi + 1:]:
        if len(r) > iv and r[im] == "gpu__time_duration.sum":
            try:
                launches.append((short(r[ik]), float(r[iv].replace(",", "")) / 1000.0))   # ns -> us
            except ValueError:
                pass
    # find the step period: distance between the last two occurrences of the NSF source kernel (once per step)
    idx = [i for i, (k, _) in enumerate(launches) if k.startswith("nsf_source")]
    if len(sys.argv) > 2:
        per = int(sys.argv[2])
    elif len(idx) >= 2:
        per = idx[-1] - idx[-2]
    else:
        per = len(launches)
    last = launches[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else launches[-per:]
    # rotate so that the step starts after the previous conv_post (end of a step)
    tot = sum(t for _, t in last)
    agg = OrderedDict()
    for k, t in last:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"{len(launches)} launches in the list; one step = {len(last)} launches, sum of kernel durations {tot / 1000.0:.3f} ms")
    print(f"{'kernel':88s} {'n':>4s} {'us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:88s} {n:4d} {t:10.1f} {100 * t / tot:6.1f}%")


if __name__ == "__main__":
    main()
