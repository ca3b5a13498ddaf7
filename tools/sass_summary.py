#!/usr/bin/env python
"""Per-kernel SASS evidence for the Blackwell-native paths: counts of tcgen05 / TMEM / TMA / mma.sync / cp.async mnemonics in every
kernel of tscd_b200/lib/libtscd_b200.so (cuobjdump -sass).   python tools/sass_summary.py > profiles/sass_summary_r2.txt

  UTCHMMA  = tcgen05.mma (f16/bf16)      LDTM / STTM = tcgen05.ld / st (TMEM)       UTCBAR = tcgen05.commit -> mbarrier
  UTMALDG  = cp.async.bulk.tensor (TMA load)       HMMA = mma.sync (small-frame kernels)      LDGSTS = cp.async (16-byte global->shared)
  SYNCS    = mbarrier try_wait / arrive"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tscd_b200", "lib", "libtscd_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "SYNCS", "MUFU.EX2", "DMUL", "DADD"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            total[cur] += 1
            for k in OPS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("arch: sm_100a   library:", os.path.relpath(LIB, ROOT))
    print(f"{'kernel':90s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in OPS))
    tot = collections.Counter()
    for (name, c), dn in zip(counts.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("tscd::", "")
        print(f"{short[:90]:90s} {total[name]:6d} " + " ".join(f"{c[k]:8d}" for k in OPS))
        tot.update(c)
    print(f"{'TOTAL':90s} {sum(total.values()):6d} " + " ".join(f"{tot[k]:8d}" for k in OPS))


if __name__ == "__main__":
    sys.exit(main())
