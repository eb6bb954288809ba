"""Per-kernel shares from an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python bench.py ...
    python tools/launch_summary.py launches.csv[.gz] > summary.md
(serialised, cold-cache launches: the SHARES are comparable with bench.py's per-launch event timing, the
absolute times are not bench numbers)."""
import collections
import csv
import gzip
import sys


def main():
    path = sys.argv[1]
    op = gzip.open if path.endswith(".gz") else open
    per = collections.defaultdict(lambda: [0, 0.0])
    with op(path, "rt") as f:
        rows = list(csv.reader(f))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
    hdr = rows[h]
    ci, cv, cu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows[h + 1:]:
        if len(r) <= cv:
            continue
        try:
            v = float(r[cv].replace(",", ""))
        except ValueError:
            continue
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[cu].lower(), 1e-3)
        name = r[ci].replace("(anonymous namespace)::", "").replace("mcgvc::", "").split("(")[0]
        per[name][0] += 1
        per[name][1] += us
    total = sum(v[1] for v in per.values())
    print("# ncu launch list summary: %s\n" % path)
    print("%d launches, %.1f ms summed kernel time (serialised, cold cache, unthrottled clocks)\n" % (sum(v[0] for v in per.values()), total / 1e3))
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---|---|---|---|")
    for name, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %.1f%% | %.1f |" % (name[:80], n, us / 1e3, 100 * us / total, us / n))


if __name__ == "__main__":
    main()
