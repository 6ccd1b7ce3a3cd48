"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown on stdout).
usage: python tools/summarize_launches.py launches.csv [title]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("<")[0].split("(")[0].replace("void ", "")
        v = float(r[-1].replace(",", ""))
        unit = r[-2]
        if unit in ("ns", "nsecond"):
            v /= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print(f"### {title}\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")
    print(f"| total | | {tot:.1f} | | |\n")


if __name__ == "__main__":
    main()
