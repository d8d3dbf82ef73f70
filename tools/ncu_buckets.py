"""Aggregates an ncu source-page dump (cuda,sass view) of the sweep kernel by source line / by kernel phase."""
import collections
import csv
import subprocess
import sys

rep, src_path = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
cols = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
src = open(src_path).read().split("\n")
# phases = function each line belongs to (nearest preceding function signature)
import re
func_at = {}
cur = "top"
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*\n)?(?:__device__|__global__).*?\b(\w+)\(", l)
    if m:
        cur = m.group(1)
    func_at[i] = cur
inst = collections.Counter(); samp = collections.Counter(); st = collections.defaultdict(collections.Counter)
linst = collections.Counter(); lsamp = collections.Counter()
curl = None
for r in rows[hi + 1:]:
    if r[0].strip().isdigit():
        curl = int(r[0]); continue
    if len(r) > 7 and r[2].startswith("0x"):
        try:
            n = int(r[7]); s = int(r[6])
        except ValueError:
            continue
        b = func_at.get(curl, "?")
        inst[b] += n; samp[b] += s; linst[curl] += n; lsamp[curl] += s
        for sn in stalls:
            try:
                st[b][sn] += int(r[cols[sn]])
            except (ValueError, IndexError):
                pass
ti = sum(inst.values()); ts = sum(samp.values())
print(f"total warp instructions {ti}, samples {ts}")
for b, n in inst.most_common():
    top = ", ".join(f"{k[6:]}:{v}" for k, v in st[b].most_common(5))
    print(f"{b:22s} inst {n / ti * 100:5.1f}%  samples {samp[b] / ts * 100:5.1f}%  | {top}")
print("--- top lines by samples")
for ln, n in lsamp.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    print(f"{ln:4d} samp {n / ts * 100:5.1f}% inst {linst[ln] / ti * 100:5.1f}% | {src[ln - 1].strip()[:100] if ln and ln <= len(src) else ''}")
