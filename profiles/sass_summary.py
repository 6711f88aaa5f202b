"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA), HMMA (legacy mma.sync: expected 0), LDGMC (multimem.ld_reduce over
NVSwitch multicast memory), plus registers per thread.

    python profiles/sass_summary.py show-and-tell_b200/libsnt_b200.so > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
demangle = lambda names: dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA", "LDGMC"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
names = demangle(list(counts))
print(f"# {so}: {len(counts)} kernels; columns = instruction counts in the SASS of each kernel (cuobjdump -sass), regs/thread, static smem")
print(f"{'UTC*MMA':>8s} {'LDTM':>5s} {'STTM':>5s} {'UTMALDG':>8s} {'UTMASTG':>8s} {'UBLKCP':>7s} {'UTCBAR':>7s} {'HMMA':>5s} {'MUFU':>5s} {'FFMA':>6s} {'LDGMC':>6s} {'insts':>6s} {'regs':>5s}  kernel")
tc = 0
for k, c in counts.items():
    n = names.get(k, k)
    n = re.sub(r"\(.*$", "", n)
    if len(n) > 150:
        n = n[:150] + "..."
    r = regs.get(k, ("?", "?"))[0]
    mma = c["UTCHMMA"] + c["UTCQMMA"]
    tc += mma > 0
    print(f"{mma:8d} {c['LDTM']:5d} {c['STTM']:5d} {c['UTMALDG']:8d} {c['UTMASTG']:8d} {c['UBLKCP']:7d} {c['UTCBAR']:7d} "
          f"{c['HMMA']:5d} {c['MUFU']:5d} {c['FFMA']:6d} {c['LDGMC']:6d} {c['_n']:6d} {str(r):>5s}  {n}")
print(f"# kernels with tcgen05.mma (UTC*MMA): {tc}; kernels with legacy HMMA: {sum(1 for c in counts.values() if c['HMMA'])}; "
      f"kernels with multimem.ld_reduce (LDGMC, NVSwitch in-flight reduction): {sum(1 for c in counts.values() if c['LDGMC'])}")
