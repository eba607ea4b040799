"""Per-kernel histogram of the SASS opcodes that tell a Blackwell-native kernel from a recompiled one
(B200_PROFILING.md): tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, tcgen05.commit = UTCBAR, bulk / tensor copies =
UBLKCP / UTMALDG / UTMASTG, legacy mma.sync = HMMA, cp.async = LDGSTS, MUFU, packed fp32 = FFMA2/FADD2/FMUL2.
usage: python tools/sass_opcodes.py [libsrwn.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sr-wavenet_b200", "libsrwn.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "MUFU", "FFMA2", "FADD2", "FMUL2",
         "SYNCS", "BAR", "RED", "ATOM", "MEMBAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
                break
print("# SASS opcode histogram of %s (cuobjdump -sass, sm_100a)" % os.path.basename(lib))
print("# kernel | instructions | " + " ".join(WATCH))
for k in counts:
    short = re.sub(r"\(.*", "", k)
    print("%-64s %6d  %s" % (short[:64], total[k], "  ".join("%s=%d" % (w, counts[k][w]) for w in WATCH if counts[k][w])))
