"""Simulated annealing over the placement (permutation) of the pair arrays of the structured HEX8 sweep so that the
twelve accumulate instructions of an element visit are (nearly) bank-conflict-free.  Row/column index m of the DMMA
tile is the node with lexicographic rank m (dx = m&1, dy = (m>>1)&1, dz = m>>2).  Output: JSON with the positions."""
import json, math, random, sys
KIJ = [(0,0),(1,1),(2,2),(0,1),(0,2),(1,2)]
INPL = {(0,0):0,(1,0):1,(-1,1):2,(0,1):3,(1,1):4}
def co(m): return (m & 1, (m >> 1) & 1, m >> 2)
def lane_access(lane, bb, k, CX):
    a, q4 = lane >> 2, lane & 3
    b = 2*q4+bb
    i, j = KIJ[k]
    if a == b: lo = hi = a; comp = k
    else:
        if a < b: lo, hi, ci, cj = a, b, i, j
        else:
            if i == j: return None
            lo, hi, ci, cj = b, a, j, i
        comp = ci*3+cj
    L, H = co(lo), co(hi)
    ox, oy, oz = H[0]-L[0], H[1]-L[1], H[2]-L[2]
    so = L[1]*CX + L[0]
    if oz == 0:
        buf = 1 if L[2] else 0
        d = INPL[(ox,oy)]
        arr = comp if d == 0 else 6 + (d-1)*9 + comp
    else:
        buf = 2; arr = ((oy+1)*3+(ox+1))*9 + comp
    return buf, arr, so
def groups(CX):
    g = []
    for bb in range(2):
        for k in range(6):
            for half in range(2):
                lst = [lane_access(l, bb, k, CX) for l in range(half*16, half*16+16)]
                g.append([x for x in lst if x])
    return g
def run(CX=9, NSP=81, iters=400000, seed=0):
    rnd = random.Random(seed)
    G = groups(CX)
    pin = list(range(42)); pcr = list(range(81)); rnd.shuffle(pin); rnd.shuffle(pcr)
    base = [0, rnd.choice((0, 8)), rnd.randrange(16)]
    def cost():
        t = 0
        for lst in G:
            banks = {}
            for buf, arr, so in lst:
                r = ((pin[arr] if buf < 2 else pcr[arr]) * NSP + base[buf] + so) % 16
                banks[r] = banks.get(r, 0) + 1
            t += max(banks.values())
        return t
    cur = cost(); best = (cur, list(pin), list(pcr), list(base))
    T = 1.0
    for it in range(iters):
        T = max(0.02, 1.0 * (1 - it / iters))
        w = rnd.randrange(10)
        if w < 4:
            i, j = rnd.randrange(42), rnd.randrange(42); pin[i], pin[j] = pin[j], pin[i]
        elif w < 9:
            i, j = rnd.randrange(81), rnd.randrange(81); pcr[i], pcr[j] = pcr[j], pcr[i]
        else:
            i = rnd.randrange(1, 3); j = base[i]; base[i] = rnd.choice((0, 8)) if i == 1 else rnd.randrange(16)
        c = cost()
        if c <= cur or rnd.random() < math.exp((cur - c) / T):
            cur = c
            if c < best[0]: best = (c, list(pin), list(pcr), list(base))
        else:
            if w < 4: pin[i], pin[j] = pin[j], pin[i]
            elif w < 9: pcr[i], pcr[j] = pcr[j], pcr[i]
            else: base[i] = j
        if best[0] == len(G): break
    return best, len(G)
if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    nsp = int(sys.argv[2]) if len(sys.argv) > 2 else 81   # slots per array (81: 9x9 tile, 63: 9x7 tile)
    best, ideal = run(seed=seed, NSP=nsp)
    print(json.dumps({"cost": best[0], "ideal": ideal, "pin": best[1], "pcr": best[2], "base": best[3]}))
