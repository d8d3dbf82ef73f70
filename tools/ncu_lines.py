"""Per-source-line warp instructions per element of an ncu report: python tools/ncu_lines.py REP SRC N_ELEM [LO HI] [MIN]"""
import collections
import csv
import subprocess
import sys

rep, srcp, nel = sys.argv[1], sys.argv[2], float(sys.argv[3])
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi_ = int(sys.argv[5]) if len(sys.argv) > 5 else 10 ** 9
mn = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
src = open(srcp).read().split("\n")
linst = collections.Counter(); lsamp = collections.Counter(); ops = collections.Counter(); cur = None
for r in rows[hi + 1:]:
    if r[0].strip().isdigit():
        cur = int(r[0]); continue
    if len(r) > 7 and r[2].startswith("0x"):
        try:
            n = int(r[7]); s = int(r[6])
        except ValueError:
            continue
        linst[cur] += n; lsamp[cur] += s
        if cur and lo <= cur <= hi_:
            t = r[3].split()
            ops[t[1] if t[0].startswith("@") and len(t) > 1 else t[0]] += n
ts = sum(lsamp.values())
tot = 0
for ln in sorted(linst):
    if lo <= ln <= hi_:
        tot += linst[ln]
        if linst[ln] / nel >= mn:
            print(ln, round(linst[ln] / nel, 1), f"{lsamp[ln] / ts * 100:.1f}%", src[ln - 1].strip()[:105] if ln <= len(src) else "")
print("total in range per element", round(tot / nel, 1))
print({k: round(v / nel, 1) for k, v in ops.most_common(28)})
