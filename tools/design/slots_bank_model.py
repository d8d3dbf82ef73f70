"""Bank model of the slot kernel (sweep_hex8_slots.cu): shared-memory wavefronts of the slot stores (element phase) and of
the gather loads per element, as a function of the array stride NEP and of the bank phase of the in-plane regions.
Banks are 8 bytes wide, 16 per wavefront; a half-warp instruction costs max over banks of the distinct addresses in it."""
import itertools, sys
KI = [0, 1, 2, 0, 0, 1]; KJ = [0, 1, 2, 1, 2, 2]
def pm(a, b): return (b & 1) * 8 + a * 2 + (b >> 1)
def updiag(d): return [0, 4, 7, 9][d]
def up(a, b): return a + updiag(b - a)

def lane_tab(lane):
    m, q4 = lane >> 2, lane & 3
    mz, bz, m4 = m >> 2, q4 >> 1, m & 3
    kind = (1 if mz == 0 else 2) if mz == bz else 0
    out = {}
    for bb in range(2):
        b4 = (2 * q4 + bb) & 3
        for k in range(6):
            idx = None
            if mz == bz:
                if k >= 3: idx = (k - 3) * 16 + pm(m4, b4)
                elif m4 <= b4: idx = 48 + k * 10 + up(m4, b4)
            elif mz == 0:
                idx = (KI[k] * 3 + KJ[k]) * 16 + pm(m4, b4)
            else:
                if k >= 3: idx = (KJ[k] * 3 + KI[k]) * 16 + pm(b4, m4)
            if idx is not None: out[bb * 6 + k] = idx
    return kind, out

def wf(addrs):
    """addrs: list of double-addresses of the active lanes of a half-warp"""
    if not addrs: return 0
    banks = {}
    for a in set(addrs): banks[a % 16] = banks.get(a % 16, 0) + 1
    return max(banks.values())

def store_cost(NEP, ph_in, ecol=0):
    tot = ideal = 0
    tabs = [lane_tab(l) for l in range(32)]
    for t in range(12):
        for half in range(2):
            st, ld = [], []
            for lane in range(half * 16, half * 16 + 16):
                kind, tab = tabs[lane]
                if t not in tab: continue
                base = 0 if kind == 0 else 100000 * 16 + ph_in   # regions far apart, phase mod 16
                a = base + tab[t] * NEP + ecol
                st.append(a)
                if kind == 1: ld.append(a)
            tot += wf(st) + wf(ld); ideal += (1 if st else 0) + (1 if ld else 0)
    return tot, ideal

def gather_cost(NEP, ph_in, nox=7, noy=7):
    DXC, DYC, DXD, DYD, CST, DST = -10 * NEP + 1, -5 * NEP + 8, -NEP + 1, -2 * NEP + 8, 16 * NEP, 10 * NEP
    n_own = nox * noy
    tot = ideal = 0
    for group in range((n_own + 9) // 10):
        for DK in (0, 1, -1):
            for DJ in (0, -1, 1):
                # instruction list: per corner, the loads; collect per instruction the addresses of the active lanes
                corners = [(cx, cy) for cy in (0, 1) for cx in (0, 1) if (cy == 0 and DJ <= 0) or (cy == 1 and DJ >= 0)]
                for cx, cy in corners:
                    ninstr = 9
                    for c in range(ninstr):
                        for half in range(2):
                            ad = []
                            for lane in range(half * 16, min(half * 16 + 16, 30)):
                                l3, di = lane // 3, lane % 3 - 1
                                col = group * 10 + l3
                                if col >= n_own: continue
                                if (cx == 0 and di > 0) or (cx == 1 and di < 0): continue
                                oyi, oxi = divmod(col, nox)
                                e00 = oyi * 8 + oxi
                                if DK == 0:
                                    base = 100000 * 16 + ph_in + e00
                                    if c < 3:
                                        a = base + ((1 + di) * 8 + 6 + (1 + DJ)) * NEP + cx * DXC + cy * DYC + c * CST
                                    elif c < 6:
                                        a = base + (8 + 2 * (1 + di) + 4 * (1 + DJ) + 1) * NEP + cx * DXC + cy * DYC + (c - 3) * CST
                                    else:
                                        dlt = di + 2 * DJ
                                        a = base + (48 + min(3, 3 + dlt) + updiag(abs(dlt))) * NEP + cx * DXD + cy * DYD + (c - 6) * DST
                                else:
                                    idx00 = (1 + di) * 8 + 6 + (1 + DJ) if DK > 0 else 8 + 2 * (1 + di) + 4 * (1 + DJ) + 1
                                    a = idx00 * NEP + e00 + cx * DXC + cy * DYC + c * CST
                                ad.append(a)
                            tot += wf(ad); ideal += 1 if ad else 0
    return tot / n_own, ideal / n_own

if __name__ == "__main__":
    for NEP in (65, 67, 69, 71):
        best = None
        for ph in range(16):
            s, si = store_cost(NEP, ph)
            g, gi = gather_cost(NEP, ph)
            obj = 1.31 * s + g
            row = (round(obj, 1), ph, s, si, round(g, 1), round(gi, 1))
            if NEP == 65 and ph == 8: print("current:", NEP, row)
            if best is None or row < best: best = row
        print("NEP", NEP, "best (objective, phase, store, store ideal, gather, gather ideal):", best)
