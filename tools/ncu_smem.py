"""Shared-memory wavefronts (actual / ideal) per element by source line: python tools/ncu_smem.py REP SRC N_ELEM"""
import collections
import csv
import subprocess
import sys

rep, srcp, nel = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
h = rows[hi]
iw, ii, ig = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal"), h.index("L2 Theoretical Sectors Global")
src = open(srcp).read().split("\n")
w = collections.Counter(); wi = collections.Counter(); g = collections.Counter(); cur = None
for r in rows[hi + 1:]:
    if r[0].strip().isdigit():
        cur = int(r[0]); continue
    if len(r) > ig and r[2].startswith("0x"):
        for c, i in ((w, iw), (wi, ii), (g, ig)):
            try:
                c[cur] += int(r[i])
            except ValueError:
                pass
print("shared wavefronts per element: total", round(sum(w.values()) / nel, 1), "ideal", round(sum(wi.values()) / nel, 1))
for ln, n in w.most_common(30):
    if n / nel >= 0.5:
        print(ln, round(n / nel, 1), "ideal", round(wi[ln] / nel, 1), "|", src[ln - 1].strip()[:100] if ln and ln <= len(src) else "")
print("global sectors per element:", round(sum(g.values()) / nel, 1))
for ln, n in g.most_common(6):
    print(ln, round(n / nel, 1), "|", src[ln - 1].strip()[:100] if ln and ln <= len(src) else "")
