"""Hot regions of one kernel from `ncu -i rep --page source --csv` (SASS view): contiguous blocks of instructions with the
same executed count (= loop bodies), their instruction counts, stall samples by reason and opcode mix.
  python tools/ncu_hot.py <source.csv> [min_exec]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ci = {k: i for i, k in enumerate(h)}
body = rows[2:]
minexec = float(sys.argv[2]) if len(sys.argv) > 2 else 0
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
blocks, cur = [], None
for n, r in enumerate(body):
    ex = float(r[ci["Instructions Executed"]] or 0)
    if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
        cur = {"ex": ex, "start": n, "n": 0, "samples": 0, "stalls": collections.Counter(), "ops": collections.Counter()}
        blocks.append(cur)
    cur["n"] += 1
    cur["samples"] += float(r[ci["# Samples"]] or 0)
    for s in stall_cols:
        v = float(r[ci[s]] or 0)
        if v:
            cur["stalls"][s[6:]] += v
    src = r[ci["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    cur["ops"][op.split(".")[0]] += 1
tot_inst = sum(b["ex"] * b["n"] for b in blocks)
tot_s = sum(b["samples"] for b in blocks)
print("total warp instructions %.3g, samples %d" % (tot_inst, tot_s))
for b in sorted(blocks, key=lambda b: -b["samples"])[:14]:
    if b["ex"] < minexec:
        continue
    print("\nblock @%d: %d instr x %.0f exec = %.1f%% of instructions, %.1f%% of samples (%.2f samples/instr-exec)" %
          (b["start"], b["n"], b["ex"], 100 * b["ex"] * b["n"] / tot_inst, 100 * b["samples"] / tot_s,
           b["samples"] / max(b["ex"] * b["n"], 1) * 1e3))
    print("   stalls: " + ", ".join("%s %.0f%%" % (k, 100 * v / max(b["samples"], 1)) for k, v in b["stalls"].most_common(6)))
    print("   ops: " + ", ".join("%s %d" % kv for kv in b["ops"].most_common(14)))
