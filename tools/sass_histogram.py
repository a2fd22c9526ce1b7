"""Opcode histogram per kernel of the built library (no GPU needed):
    python tools/sass_histogram.py pddp_b200/lib/libpddp_b200.so > profiles/r2_sass_histogram.md
Lists, per kernel, the instruction count and the opcodes that show which hardware paths the kernel uses
(tcgen05: UTCHMMA / UTCBAR / LDTM, bulk copies: UBLKCP, tensor-map TMA: UTMALDG / UTMASTG, mbarrier: SYNCS,
packed fp32: FFMA2, register re-balancing: USETMAXREG, ...) and the ten most frequent opcodes."""
import collections
import re
import subprocess
import sys

MARKERS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ELECT", "FFMA2",
           "USETMAXREG", "LDGSTS", "DFMA", "MUFU", "HMMA", "REDUX", "BAR", "ATOMG", "RED")


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("pddp::", "")
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and name:
            op = m.group(1)
            if op == "UTCHMMA" and m.group(2) and "2CTA" in m.group(2):
                op = "UTCHMMA.2CTA"
            kernels[name][op] += 1
    print("# SASS opcode histogram per kernel (`cuobjdump -sass %s`)\n" % path)
    print("| kernel | instructions | hardware-path opcodes | most frequent opcodes |")
    print("|---|---|---|---|")
    for k, c in sorted(kernels.items(), key=lambda kv: -sum(kv[1].values())):
        marks = ", ".join("%s x%d" % (op, n) for op, n in c.items() if any(op.startswith(mk) for mk in MARKERS))
        top = ", ".join("%s %d" % (op, n) for op, n in c.most_common(10))
        print("| `%s` | %d | %s | %s |" % (k, sum(c.values()), marks or "-", top))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "pddp_b200/lib/libpddp_b200.so")
