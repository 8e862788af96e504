"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and mean time.

    python tools/ncu_launches.py gpurun_out/launches.csv [substring] [--last N]
"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    recs = [(re.sub(r"\(.*", "", r[4]).replace("void ", ""), float(r[-1])) for r in rows[hdr + 1:] if len(r) > 5 and sub in r[4]]
    if last:
        recs = recs[-last:]
        for n, t in recs:
            print("%-60s %10.1f us" % (n[:60], t / 1e3))
        return
    agg = {}
    for n, t in recs:
        c, s = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, s + t)
    tot = sum(s for _, s in agg.values())
    for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s n=%5d total %10.1f us  mean %9.1f us  share %5.1f%%" % (n[:60], c, s / 1e3, s / c / 1e3, 100 * s / tot))


if __name__ == "__main__":
    main()
