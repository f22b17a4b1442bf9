"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

    python profiles/summarize_launches.py gpurun_out/launches.csv [--last N] > profiles/<name>.md

`--last N` keeps only the last N launches (e.g. one bench step). Times under ncu are cold-cache and
serialised: compare SHARES, not absolutes (B200_PROFILING.md).
"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=0)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    with open(args.csv) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    if args.last:
        rows = rows[-args.last:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"]))[:80]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    print(f"launches: {len(rows)}   sum of kernel durations: {total / 1e3:.3f} ms\n")
    print("| us | launches | share | kernel |")
    print("|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print(f"| {v[1]:.1f} | {v[0]} | {100 * v[1] / total:.1f}% | `{k}` |")


if __name__ == "__main__":
    main()
