"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean,
share of the profiled window (kernels keyed by name + grid so the GEMM shapes stay apart).
usage: python profiles/summarize_launches.py <csv> [skip_first_n]"""
import collections
import csv
import sys


def main(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, ""])
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        name, val = row.get("Kernel Name"), row.get("Metric Value")
        if not name or not val:
            continue
        if skip > 0:
            skip -= 1
            continue
        v = float(val.replace(",", "")) / 1000.0
        k = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0] + " " + row["Grid Size"].replace(" ", "")
        agg[k][0] += 1
        agg[k][1] += v
        agg[k][2] = row["Grid Size"] + row["Block Size"]
        tot += v
        n += 1
    print("%-48s %5s %10s %9s %7s  %s" % ("kernel", "n", "total_us", "avg_us", "share", "grid/block"))
    for k, (c, t, g) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-48s %5d %10.1f %9.2f %6.1f%%  %s" % (k[:48], c, t, t / c, 100 * t / tot, g))
    print("total %.1f us over %d launches" % (tot, n))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
