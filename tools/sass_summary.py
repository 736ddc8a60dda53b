"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (no GPU needed):
  python tools/sass_summary.py > profiles/sass_summary.txt
UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG/UTMAREDG = TMA load / store /
reduce-add, UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS = cp.async, DFMA/DMUL/DADD = FP64 pipe, MUFU = SFU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "htr-vt_b200", "libhtrvt_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "LDGSTS", "DFMA",
       "MUFU", "ATOMS", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name, cnt, size = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("htrvt::", "")
            cnt[name] = collections.Counter()
            size[name] = 0
            continue
        if name is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            size[name] += 1
            op = m.group(1)
            for p in PAT:
                if op.startswith(p):
                    cnt[name][p] += 1
    print("library: htr-vt_b200/libhtrvt_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a), %d kernels" % len(cnt))
    print("%-64s %7s " % ("kernel", "SASS") + " ".join("%8s" % p for p in PAT))
    tot = collections.Counter()
    for k in sorted(cnt):
        tot.update(cnt[k])
        print("%-64s %7d " % (k[:64], size[k]) + " ".join("%8d" % cnt[k][p] for p in PAT))
    print("%-64s %7d " % ("TOTAL", sum(size.values())) + " ".join("%8d" % tot[p] for p in PAT))


if __name__ == "__main__":
    sys.exit(main())
