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


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
