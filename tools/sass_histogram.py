"""SASS opcode histogram per kernel of the built library: python tools/sass_histogram.py > profiles/rNN_sass_histogram.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mast-multiphysics_b200", "libmast_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur = None
hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
dem = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
special = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DMMA", "LDGSTS", "UTCMMA", "FENCE", "LDL", "STL")
print("# SASS opcode histograms of the built library (cuobjdump -sass libmast_b200.so, sm_100a), kernels over 300 instructions\n")
print("Static instruction counts per kernel (opcode family = mnemonic up to the first dot).  `DMMA` = `mma.sync.m8n8k4.f64`; `UBLKCP` ="
      " `cp.async.bulk` (TMA bulk copy: the accumulator clear and the nv = 6 rotational rows of `hex8_struct_kernel`), `SYNCS` = mbarrier"
      " operations, `LDGSTS` = `cp.async` staging of the node planes, `LDL` / `STL` = local-memory traffic (spills / indexed locals)."
      "  No `UTCMMA`: tcgen05 has no fp64.\n")
rows = []
for mangled, name in zip(hist, dem):
    h = hist[mangled]
    n = sum(h.values())
    if n < 300: continue
    short = re.sub(r"mastb200::\(anonymous namespace\)::", "", name)
    short = re.sub(r"\(.*", "", short)
    top = ", ".join(f"{k} {v}" for k, v in h.most_common(12))
    sp = ", ".join(f"{k} {h[k]}" for k in special if h.get(k))
    rows.append((short, n, top, sp))
for short, n, top, sp in sorted(rows):
    print(f"* `{short}` -- {n} instructions: {top}" + (f"  |  **{sp}**" if sp else ""))
