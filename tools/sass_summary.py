"""SASS evidence that the contractions are Blackwell-native: per-kernel counts of the tcgen05 / TMEM / TMA mnemonics in
libvqa_b200.so (cuobjdump -sass), written to profiles/<round>_sass_summary.txt.

    python tools/sass_summary.py [profiles/r2_sass_summary.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "t5-resnet-vqa_b200", "libvqa_b200.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS",
             "HMMA", "IMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            counts[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["instructions"] += 1
        base = op.split(".")[0]
        if base in MNEMONICS:
            counts[cur][base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                counts[cur]["UTCHMMA.2CTA"] += 1
    names = demangle(list(counts))
    total = collections.Counter()
    rows = []
    for k, c in counts.items():
        for m_ in MNEMONICS:
            total[m_] += c[m_]
        short = names.get(k, k).replace("vqa::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(.*", "", short)
        rows.append((short, c))
    with open(dst, "w") as f:
        f.write("# cuobjdump -sass %s (sm_100a): tcgen05 / TMEM / TMA mnemonics per kernel\n" % os.path.relpath(LIB, ROOT))
        f.write("# UTCHMMA = tcgen05.mma (kind::f16), .2CTA = cta_group::2; LDTM/STTM = tcgen05.ld/st (TMEM); UTMALDG/UTMASTG/UTMAREDG "
                "= TMA tensor load / store / reduce; HMMA/IMMA = legacy mma.sync (must be 0)\n")
        f.write("totals: " + ", ".join("%s %d" % (m_, total[m_]) for m_ in MNEMONICS) + "\n\n")
        f.write("%-78s %6s %8s %6s %5s %5s %8s %8s %9s %5s\n" % ("kernel", "instr", "UTCHMMA", ".2CTA", "LDTM", "STTM", "UTMALDG",
                                                                "UTMASTG", "UTMAREDG", "HMMA"))
        for short, c in sorted(rows, key=lambda r: (-r[1]["UTCHMMA"], r[0])):
            f.write("%-78s %6d %8d %6d %5d %5d %8d %8d %9d %5d\n" % (short[:78], c["instructions"], c["UTCHMMA"], c["UTCHMMA.2CTA"],
                                                                    c["LDTM"], c["STTM"], c["UTMALDG"], c["UTMASTG"],
                                                                    c["UTMAREDG"], c["HMMA"]))
    print("wrote", dst, dict(total))


if __name__ == "__main__":
    main()
