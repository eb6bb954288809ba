"""Per-launch summary of an `ncu --set full` capture:
    ncu -i <rep>.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv[.gz] [title] > summary.md
Columns: duration, tensor-pipe activity, DRAM bytes, achieved HBM GB/s against the pool's measured copy
bandwidth (MEASURED_PEAKS.json), L2 hit rate, registers, grid / cluster."""
import csv
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(s):
    try:
        return float(str(s).replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    json_out = None
    if "--json" in sys.argv:          # also write {kernel: {launches, dram_bytes_per_launch, avg_us}} for bench.py's roofline.traffic
        i = sys.argv.index("--json")
        json_out = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(path)
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        rows = list(csv.reader(f))
    # the csv may start with ncu banner lines: find the header
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[h], rows[h + 1], rows[h + 2:]
    col = {n: i for i, n in enumerate(hdr)}

    def find(sub):
        for n, i in col.items():
            if sub in n:
                return i
        return None

    c_name, c_grid = col["Kernel Name"], col.get("Grid Size")
    c_dur = col.get("gpu__time_duration.sum")
    c_tensor = find("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
    if c_tensor is None:
        c_tensor = find("sm__inst_executed_pipe_tensor")
    c_tmem = col.get("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    c_rd, c_wr = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
    c_hit = find("lts__t_sector_hit_rate.pct")
    c_regs = find("launch__registers_per_thread")
    c_clu = find("launch__cluster_dim_x")
    c_smem = find("launch__shared_mem_per_block_dynamic")
    peaks = {"hbm_gbs": 6448.4}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peaks = json.load(open(pp))
    hbm = float(peaks.get("hbm_gbs", 6448.4))

    def scale(ci, v):    # bytes / time units differ between ncu versions
        u = units[ci].lower() if ci is not None and ci < len(units) else ""
        if u in ("mbyte", "mb"):
            return v * 1e6
        if u in ("kbyte", "kb"):
            return v * 1e3
        if u in ("gbyte", "gb"):
            return v * 1e9
        return v

    def dur_us(v):
        u = units[c_dur].lower()
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)

    print("# %s\n" % title)
    print("HBM peak used for the fraction: %.1f GB/s (MEASURED_PEAKS.json copy bandwidth).  Launches are isolated, cold-cache, "
          "unthrottled clock: durations are NOT bench numbers.\n" % hbm)
    print("tensor pipe % = sm__pipe_tensor_cycles_active_realtime (% of peak sustained, elapsed; the recipe's metric); tensor busy % = "
          "sm__mem_tensor_cycles_active (% of peak sustained, active cycles: the round-1 summaries' column, tracks issued-MMA utilisation).\n")
    print("| # | kernel | grid | cluster | regs | dyn smem KB | time us | tensor pipe % | tensor busy % | DRAM rd MB | DRAM wr MB | HBM GB/s | % of HBM peak | L2 hit % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    agg = {}
    for i, r in enumerate(data):
        if len(r) <= c_name:
            continue
        name = r[c_name].replace("(anonymous namespace)::", "").replace("mcgvc::", "")
        name = name.split("(")[0]
        t = dur_us(num(r[c_dur]))
        rd = scale(c_rd, num(r[c_rd])) if c_rd is not None else float("nan")
        wr = scale(c_wr, num(r[c_wr])) if c_wr is not None else float("nan")
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else float("nan")
        tp = num(r[c_tensor]) if c_tensor is not None else float("nan")
        hit = num(r[c_hit]) if c_hit is not None else float("nan")
        tm = num(r[c_tmem]) if c_tmem is not None else float("nan")
        print("| %d | `%s` | %s | %s | %s | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f |" % (
            i, name[:60], r[c_grid] if c_grid is not None else "", r[c_clu] if c_clu is not None else "",
            r[c_regs] if c_regs is not None else "", scale(c_smem, num(r[c_smem])) / 1024 if c_smem is not None else float("nan"),
            t, tp, tm, rd / 1e6, wr / 1e6, gbs, 100 * gbs / hbm, hit))
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += rd + wr; a[3] += tp * t; a[4] = max(a[4], 100 * gbs / hbm); a[5] += (tm if tm == tm else 0.0) * t
    print("\n## per kernel\n")
    print("| kernel | launches | total us | avg us | time-weighted tensor pipe % | time-weighted tensor busy % | avg HBM GB/s | best % of HBM peak | DRAM bytes per launch MB |")
    print("|---|---|---|---|---|---|---|---|---|")
    if json_out:
        import subprocess
        try:
            commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        except Exception:
            commit = ""
        js = {"source": os.path.basename(path), "title": title, "summarised_at_commit": commit,
              "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the kernel's own launches in the capture "
                      "(ncu --set full, isolated cold-cache launches)",
              "kernels": {name.replace("void ", "").replace("<unnamed>::", ""): {"launches": n, "dram_bytes_per_launch": by / n, "avg_us": t / n}
                          for name, (n, t, by, tpw, best, tmw) in agg.items()}}
        with open(json_out, "w") as f:
            json.dump(js, f, indent=1)
    for name, (n, t, by, tpw, best, tmw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f |" % (name[:60], n, t, t / n, tpw / t if t else 0, tmw / t if t else 0,
                                                                        by / (t * 1e-6) / 1e9 if t else 0, best, by / n / 1e6))


if __name__ == "__main__":
    main()
