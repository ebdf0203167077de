"""Counts of the SASS mnemonics that show what the built library's kernels are made of (no GPU needed):
    python tools/sass_excerpt.py > profiles/<tag>_sass_excerpt.txt
UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, RED / REDG = row updates, ATOMS = shared-memory combining,
LDG.E.*.NA / .CONSTANT = the streaming matrix loads, LDG.E.STRONG = the acquire polls, MEMBAR / CCTL = fences."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hisparse_b200", "libhisparse_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UBLKCP[.\w]*|SYNCS[.\w]*|REDG?[.\w]*|ATOMS[.\w]*|ATOMG[.\w]*|LDG\.E[.\w]*|LDS[.\w]*|STG[.\w]*|MEMBAR[.\w]*|CCTL[.\w]*|ACQBULK|"
                 r"IMAD\.WIDE\.U32|VIMNMX[.\w]*|FMUL|FADD|SHFL[.\w]*|VOTE[.\w]*|BAR[.\w]*|ERRBAR|NANOSLEEP[.\w]*)\b")
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"hsb::\(anonymous namespace\)::|\(anonymous namespace\)::", "", fn)
        counts[fn] = collections.Counter()
        continue
    if fn and "/*" in line:
        body = re.sub(r"/\*[0-9a-fx]+\*/", "", line)
        for mm in pat.findall(body):
            counts[fn][mm] += 1
for fn, c in counts.items():
    if not any(k in fn for k in ("spmv_tiles_kernel", "spmv_iterate_kernel", "drain_kernel", "axpb", "wait_flags", "k_decode", "k_fill")):
        continue
    print(fn[:150])
    print("   " + "  ".join("%s x%d" % kv for kv in sorted(c.items(), key=lambda kv: (-kv[1], kv[0]))))
