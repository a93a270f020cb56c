"""Opcode census of the shipped library (`cuobjdump -sass`): per kernel, the instruction count and the tcgen05 / TMEM / TMA /
mbarrier opcodes that prove which hardware path a kernel uses (UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld,
UTMALDG / UTMASTG = TMA tensor loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UTCATOMSWS = TMEM alloc).
usage: python tools/sass_opcodes.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "lungmask_b200/liblungmask_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEY = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UTMACMDFLUSH", "UTCBAR", "UTCATOMSWS", "SYNCS", "UBLKCP",
       "HMMA", "IMMA", "FENCE", "MEMBAR", "UCGABAR", "ATOMG", "REDG", "ATOMS", "RED", "ATOM", "LDGSTS", "SHFL", "VOTE", "MATCH", "BAR")
fn, ops = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        ops[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and fn:
        ops[fn][m.group(1)] += 1
print("# %s: SASS opcode census (sm_100a), %d kernels" % (lib, len(ops)))
tot = collections.Counter()
for fn, c in ops.items():
    base = collections.Counter()
    for op, n in c.items():
        base[op.split(".")[0]] += n
    keys = {k: v for k, v in base.items() if k in KEY}
    name = re.sub(r"\(.*", "", re.sub(r"\((int|bool|unsigned int)\)", "", demangle(fn))).replace("lm::(anonymous namespace)::", "").replace("void ", "")
    print("\n%s  [%d instructions]" % (name, sum(c.values())))
    print("   " + ", ".join("%s %d" % (k, v) for k, v in sorted(keys.items(), key=lambda kv: -kv[1])))
    det = {op: n for op, n in c.items() if op.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTCATOMSWS")}
    if det:
        print("   detail: " + ", ".join("%s x%d" % (k, v) for k, v in sorted(det.items())))
    tot.update(keys)
print("\n# library totals: " + ", ".join("%s %d" % (k, v) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])))
