"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv ...`):

  python tools/launch_summary.py profiles/r2_launches_bench.csv > profiles/r2_launches_bench_summary.csv

Kernel names are cut at the argument list; shares are of the listed launches (cold-cache, serialised: compare shares,
not absolute times)."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) >= 15]
    head = next(i for i, r in enumerate(rows) if r[0] == "ID")
    name_i, val_i, unit_i = rows[head].index("Kernel Name"), rows[head].index("Metric Value"), rows[head].index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[head + 1:]:
        name = r[name_i].split("(")[0].replace("hf::", "")
        ns = float(r[val_i].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[unit_i], 1.0)
        tot[name] += ns
        cnt[name] += 1
    whole = sum(tot.values()) or 1.0
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct"])
    for k in sorted(tot, key=tot.get, reverse=True):
        w.writerow([k, cnt[k], f"{tot[k] / 1e3:.1f}", f"{tot[k] / 1e3 / cnt[k]:.2f}", f"{100 * tot[k] / whole:.1f}"])


if __name__ == "__main__":
    main(sys.argv[1])
