"""SASS opcode histogram of the kernels in a compiled object (no GPU needed):
   python tools/sass_histogram.py scuff-em_b200/csrc/far_kernel.o [name-regex] > profiles/X.sass.txt
Per kernel: instruction count, the opcodes that prove the design points (DFMA/DMUL/DADD = FP64 pipe, UBLKCP/UBLKRED = TMA bulk
copy / bulk reduction, SYNCS = mbarrier, RED/ATOM = L2 atomics, LDS/STS = shared memory, MUFU = SFU seed) and the top 12 opcodes."""
import collections, re, subprocess, sys
obj = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name = None; hist = {}
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[name] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        hist[name][m.group(1).split(".")[0]] += 1
KEY = ["DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "ATOMS", "UBLKCP", "UBLKRED", "SYNCS", "SHFL", "BAR", "FSEL", "ISETP", "IMAD", "LDC", "LDCU", "BRA", "CALL"]
for k, h in hist.items():
    if pat and not pat.search(k):
        continue
    tot = sum(h.values())
    print("== %s: %d SASS instructions" % (k, tot))
    print("   " + "  ".join("%s %d" % (o, h[o]) for o in KEY if h[o]))
    print("   top: " + "  ".join("%s %d" % (o, c) for o, c in h.most_common(12)))
