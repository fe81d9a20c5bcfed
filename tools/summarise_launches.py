"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the captured window)."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).split("::")[-1]
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        us = v / 1000 if r[mu] in ("ns", "nsecond") else (v if r[mu] in ("us", "usecond") else v * 1000)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1000:.3f} ms summed device time "
          f"(cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'share':>7} {'total_us':>11} {'launches':>9} {'avg_us':>9}  kernel")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * us / tot:6.1f}% {us:11.1f} {n:9d} {us / n:9.2f}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
