"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_summary.py <csv> <n last launches> <out>
Per kernel name: launches, total ms, share of the listed launches (cold-cache, serialised: compare SHARES, B200_PROFILING.md)."""
import csv
import sys


def main():
    path, last, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    if last > 0:
        rows = rows[-last:]
    agg = {}
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1e-6)
        name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none: last %d launches of %s\n" % (len(rows), path))
        f.write("%-60s %8s %12s %8s\n" % ("kernel", "launches", "ms", "share"))
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.3f %7.2f%%\n" % (k[:60], n, ms, 100.0 * ms / tot))
        f.write("%-60s %8d %12.3f\n" % ("total", len(rows), tot))
    print(open(out).read())


if __name__ == "__main__":
    main()
