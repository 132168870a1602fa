#!/usr/bin/env python
"""Trim an `ncu --csv --metrics gpu__time_duration.sum` launch list to ONE step of bench.py (the first complete step after the
warm-up: from one launch of the first kernel of a step up to the next) and print the per-kernel-family shares.
usage: tools/last_step.py <all_launches.csv> <out.csv>"""
import csv
import sys
from collections import defaultdict


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    head, body = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    k, v, m = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Name")
    body = [r for r in body if r[m] == "gpu__time_duration.sum"]
    # a step starts with the 3x3 stem conv on the uint8 image (with the separate normalisation kernel in older captures)
    first = "tc_stem_u8_kernel<3" if any("tc_stem_u8_kernel<3" in r[k] for r in body) else "preprocess_u8"
    starts = [i for i, r in enumerate(body) if first in r[k]]
    if len(starts) < 5:
        raise SystemExit(f"expected >= 5 steps (3 warm-up + 1 timed + e2e) in {src}, found {len(starts)}")
    step = body[starts[3]:starts[4]]           # the timed step
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(head)
        w.writerows(step)
    fam = defaultdict(lambda: [0, 0.0])
    for r in step:
        name = r[k].split("(")[0].split("<")[0].replace("kg::", "")
        fam[name][0] += 1
        fam[name][1] += float(r[v].replace(",", ""))
    total = sum(t for _, t in fam.values())
    unit = head.index("Metric Unit")
    print(f"{len(step)} launches, {total:.0f} {step[0][unit]} serialised")
    for name, (n, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:40s} {n:4d} launches  {t:12.0f}  {100 * t / total:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
