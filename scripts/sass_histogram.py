#!/usr/bin/env python
"""Per-kernel SASS instruction histogram of librefil_b200.so (cuobjdump -sass; runs without a GPU): the Blackwell-native
mnemonics -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA tensor load / store /
reduce-add), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier) -- next to the legacy tensor path (HMMA) and the
fp32 pipe (FFMA).

    python scripts/sass_histogram.py [--md] > profiles/sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "refil_b200", "librefil_b200.so")
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "LDS", "STS",
         "LDG", "STG", "ATOMG", "RED", "SHFL", "MUFU", "BAR"]


def histogram(so=SO):
    """-> {demangled kernel name: Counter(mnemonic prefix -> count)}"""
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for wname in WATCH:
                if op == wname or op.startswith(wname + "."):
                    cur[wname] += 1
    names = list(kernels)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    if len(dem) != len(names):
        dem = names
    def short(d):
        d = d.replace("void ", "").replace("(int)", "").replace("(bool)", "")
        return re.sub(r"\(.*", "", d)
    return collections.OrderedDict((short(d), kernels[n]) for d, n in zip(dem, names))


def main():
    h = histogram()
    cols = [w for w in WATCH if any(c[w] for c in h.values())]
    print("# SASS instruction histogram per kernel (`python scripts/sass_histogram.py`, cuobjdump -sass of librefil_b200.so, sm_100a)\n")
    print("| kernel | instructions | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, c in sorted(h.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 10 ** 6 - kv[1]["_total"]):
        print("| `%s` | %d | " % (k[:70], c["_total"]) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")


if __name__ == "__main__":
    main()
