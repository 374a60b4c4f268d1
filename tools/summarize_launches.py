"""Summarises an ncu --csv launch list (gpu__time_duration.sum) per kernel name: count, total us, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = r["Kernel Name"]
        rows.append((name, us))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        n = re.sub(r"\(.*", "", n)
        agg[n][0] += 1
        agg[n][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms total (serialised, cold cache)")
    print(f"{'share':>6} {'total_us':>10} {'count':>6} {'avg_us':>9}  kernel")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * us / total:6.2f} {us:10.1f} {c:6d} {us / c:9.2f}  {n[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
