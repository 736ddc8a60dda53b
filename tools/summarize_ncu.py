"""Summarise ncu outputs into markdown (run in the dev container, no GPU needed).
  python tools/summarize_ncu.py launches <launches.csv>             per-kernel launch count / time / share
  python tools/summarize_ncu.py raw <report.ncu-rep> [regex]        key counters per captured launch
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("htrvt::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("| kernel | launches | sum us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("| `%s` | %d | %.1f | %.1f%% |" % (k[:72], n, t, 100 * t / tot))
    print("\nTotal %.1f us in %d launches." % (tot, sum(a[0] for a in agg.values())))


def raw(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size"]
    idx = [hdr.index(w) for w in want if w in hdr]
    names = [w for w in want if w in hdr]
    ki = hdr.index("Kernel Name")
    print("units (row 2 of the ncu export): " + ", ".join("%s [%s]" % (n.split(".")[0], rows[1][i]) for n, i in zip(names, idx)))
    print("\n| kernel | " + " | ".join(n.split(".")[0].replace("__", " ") for n in names) + " |")
    print("|---|" + "---:|" * len(names))
    for r in rows[2:]:
        n = re.sub(r"\(CUtensor.*|\(const.*|\(htrvt.*", "", r[ki]).replace("void htrvt::", "").replace("htrvt::", "")
        if pattern and not re.search(pattern, n):
            continue
        print("| `%s` | %s |" % (n[:60], " | ".join(r[i] for i in idx)))


def counters(path, trimmed_out=None):
    """Aggregate the `--page raw --csv` export of tools/ncu_step.sh's metrics pass per kernel name."""
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    units = rows[start]
    col = {h: i for i, h in enumerate(hdr)}
    keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__cluster_size", "sm__cycles_elapsed.avg.per_second"]
    keep = [k for k in keep if k in col]

    def num(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except Exception:
            return 0.0

    def scale(k, want):           # ncu picks units per column: normalise bytes to MB and time to us
        u = units[col[k]].lower()
        f = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
             "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
        return f.get(u, 1.0)

    body = [r for r in rows[start + 1:] if len(r) > col["gpu__time_duration.sum"]]
    if trimmed_out:
        with open(trimmed_out, "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(keep)
            w.writerow([units[col[k]] for k in keep])
            for r in body:
                w.writerow([r[col[k]] for k in keep])
    agg = collections.OrderedDict()
    ts, ds = scale("gpu__time_duration.sum", "us"), scale("dram__bytes_read.sum", "MB")
    dws = scale("dram__bytes_write.sum", "MB")
    tot = 0.0
    for r in body:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("htrvt::", "")
        t = num(r, "gpu__time_duration.sum") * ts
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += num(r, "dram__bytes_read.sum") * ds + num(r, "dram__bytes_write.sum") * dws
        a[3] += t * num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        a[4] += t * num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        a[5] += t * num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
        tot += t
    print("| kernel | launches | sum us | share | DRAM MB (r+w) | DRAM GB/s | tensor pipe % | dram % of peak | L2 % |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:34]:
        n, t, mb, tp, dp, lp = a
        print("| `%s` | %d | %.1f | %.1f%% | %.1f | %.0f | %.1f | %.1f | %.1f |"
              % (k[:60], n, t, 100 * t / tot, mb, mb / t * 1e3 if t else 0, tp / t if t else 0, dp / t if t else 0,
                 lp / t if t else 0))
    print("\nTotal %.1f us in %d launches (per-launch times under ncu are serialised and cold-cache)."
          % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "counters":
        counters(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
