"""SASS opcode summary of libtlb200.so: per kernel family, how many tensor-core / TMA / TMEM / DMMA instructions the
sm_100a code holds (cuobjdump -sass).  Evidence that the hot kernels are tcgen05 + TMA code, not recompiled mma.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tensorly_b200", "csrc", "libtlb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "DMMA", "HMMA", "DFMA",
       "FFMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "ERRBAR", "ELECT", "F2FP"]
fn = None
counts = collections.OrderedDict()
arch = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = fn.replace("(anonymous namespace)::", "").replace("tlb200::", "").replace("void ", "")
        counts[fn] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if fn is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[fn]["total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[fn][o] += 1
print(f"# {os.path.relpath(so, ROOT)}: {len(counts)} kernels, arch {arch}")
print(f"# columns: instructions per kernel; only kernels with tensor-core / TMA / TMEM / DMMA opcodes are listed in full")
hdr = ["UTCHMMA", "UTCBAR", "UTMALDG", "LDTM", "STTM", "SYNCS", "DMMA", "LDGSTS", "F2FP", "total"]
print(f"{'kernel':90s} " + " ".join(f"{h:>8s}" for h in hdr))
rest = 0
for fn, c in counts.items():
    if any(c[o] for o in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "DMMA", "LDGSTS")):
        print(f"{fn[:90]:90s} " + " ".join(f"{c[h]:8d}" for h in hdr))
    else:
        rest += 1
print(f"# {rest} other kernels (SIMT: layout copies, Khatri-Rao, solves, reductions, HALS, reconstruction, comm)")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# library totals: " + ", ".join(f"{o}={tot[o]}" for o in OPS if tot[o]))
if tot["HMMA"]:
    print("# note: HMMA present (mma.sync path)", file=sys.stderr)
