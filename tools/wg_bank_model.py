#!/usr/bin/env python3
"""Shared-memory bank model of the warp-group kernels (chinium_b200/csrc/eri_wg.cuh) and search for the scratch strides.

The ncu captures of round 2 (profiles/r2d_c18_wg_ncu_full_summary.txt) show the warp-group kernels at 65-76 % of the L1 data
pipe with 35-47 % of the shared-memory wavefronts being bank-conflict replays: they are bound by shared-memory wavefronts,
not by the FP64 pipe.  This script replays the address pattern of phase B (VRR columns stored with STS.128) and phase C
(columns loaded with LDS.128) of one primitive quartet per warp, counts wavefronts per quarter-warp
(8 lanes x 16 B; distinct addresses in the same 16-byte bank group serialise, equal addresses broadcast) and searches the
three strides (column LABP, (root, direction) block TSZ, quartet scratch SCR) for the minimum.

    python tools/wg_bank_model.py            # table: current strides vs best
    python tools/wg_bank_model.py --emit     # C++ switch for eri_wg.cuh (wg_strides)
"""
import itertools
import sys

CLS = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (3, 2), (3, 3)]
NAMES = "spdf"

# key -> (MK, swap, HS, MINB)   (wg_cfg_base of eri_wg.cuh)
CFG = {42: (3, 0, 1, 2), 44: (3, 0, 1, 2), 52: (1, 0, 1, 2), 53: (1, 0, 1, 2), 54: (2, 0, 1, 2), 55: (2, 0, 1, 2),
       64: (3, 1, 1, 2), 65: (1, 1, 1, 3), 72: (1, 0, 1, 3), 73: (1, 0, 1, 3), 74: (2, 0, 1, 2), 75: (2, 1, 1, 2),
       76: (2, 0, 1, 2), 77: (2, 0, 1, 2), 81: (1, 0, 1, 2), 82: (1, 0, 1, 2), 83: (1, 0, 1, 2), 84: (4, 1, 1, 2),
       85: (2, 1, 1, 2), 86: (1, 0, 1, 2), 87: (1, 0, 1, 2), 88: (2, 0, 2, 2), 91: (1, 0, 2, 2), 92: (1, 0, 2, 2),
       93: (1, 0, 2, 2), 94: (4, 1, 1, 2), 95: (4, 1, 2, 2), 96: (1, 0, 2, 2), 97: (1, 0, 2, 2), 98: (2, 0, 5, 2),
       99: (4, 0, 10, 2), 71: (1, 0, 1, 3), 51: (1, 0, 1, 2), 66: (1, 0, 1, 4), 43: (3, 0, 1, 2)}


# alternatives measured as variant 1 (wg_cfg_alt1): the model's pick where it differs from the base configuration
ALT = {84: (1, 0, 1, 2), 94: (1, 0, 2, 2), 97: (1, 0, 2, 2), 53: (2, 0, 1, 2), 73: (2, 0, 1, 2), 74: (1, 0, 1, 2), 54: (1, 0, 1, 2),
       43: (3, 0, 1, 2), 66: (3, 0, 1, 2), 76: (1, 0, 1, 2), 75: (1, 1, 1, 2), 77: (1, 0, 1, 2), 44: (1, 0, 1, 2)}
for _k, _v in ALT.items():
    CFG[100 + _k] = _v


def ncart(l):
    return (l + 1) * (l + 2) // 2


def cart_row(n):
    return 3 if n >= 6 else 2 if n >= 3 else 1 if n >= 1 else 0


def cart_exps(l, n):
    r = cart_row(n)
    lz = n - r * (r + 1) // 2
    return (l - r, r - lz, lz)


def wg_pad(n):
    return n + ((6 - n % 4) % 4)


def shape(key):
    mk, swap, hs, minb = CFG[key]
    bra, ket = CLS[(key % 100) // 10], CLS[key % 10]
    (la, lb), (lc, ld) = (ket, bra) if swap else (bra, ket)
    return la, lb, lc, ld, mk, hs, minb


def sizes(la, lb, lc, ld, mk, hs):
    na, nb, nc, nd = ncart(la), ncart(lb), ncart(lc), ncart(ld)
    ncd = nc * nd
    nroots = (la + lb + lc + ld) // 2 + 1
    gs = (ncd + mk - 1) // mk
    qw = 32 // gs
    nap = na // hs
    ne = nap * nb
    return dict(na=na, nb=nb, nc=nc, nd=nd, ncd=ncd, nroots=nroots, gs=gs, qw=qw, nap=nap, ne=ne, lab=la + lb,
                r1=ne * gs, r2=(nap + nb) * ncd, rwp=(2 * nroots + 1) & ~1)


def current_strides(la, lb, lc, ld, mk, hs):
    s = sizes(la, lb, lc, ld, mk, hs)
    labp = wg_pad(s["lab"] + 1)
    tsz = wg_pad((lc + 1) * (ld + 1) * labp)
    tq = s["nroots"] * 3 * tsz + s["rwp"]
    scr = wg_pad(max(tq, s["r1"], s["r2"]))
    return labp, tsz, scr


def wavefronts128(addrs):
    """addrs: list of 32 entries (double index, even) or None.  Quarter-warp phases."""
    tot = 0
    for qw in range(4):
        groups = {}
        for a in addrs[8 * qw:8 * qw + 8]:
            if a is None:
                continue
            groups.setdefault((a // 2) % 8, set()).add(a)
        if groups:
            tot += max(len(v) for v in groups.values())
    return tot


def wavefronts64(addrs):
    tot = 0
    for hw in range(2):
        groups = {}
        for a in addrs[16 * hw:16 * hw + 16]:
            if a is None:
                continue
            groups.setdefault(a % 16, set()).add(a)
        if groups:
            tot += max(len(v) for v in groups.values())
    return tot


def simulate(la, lb, lc, ld, mk, hs, labp, tsz, scr):
    s = sizes(la, lb, lc, ld, mk, hs)
    gs, qw, nroots, lab, ncd = s["gs"], s["qw"], s["nroots"], s["lab"], s["ncd"]
    lanes = [(l // gs, l % gs) if l // gs < qw else None for l in range(32)]
    wB = wC = idealB = idealC = 0
    # phase B: stores
    npass = (3 * nroots + gs - 1) // gs
    for p in range(npass):
        for l_ in range(ld + 1):
            for k in range(lc + 1):
                col = (k * (ld + 1) + l_) * labp
                nchunk = (lab + 1) // 2
                for ch in range(nchunk):
                    ad = []
                    for ln in lanes:
                        if ln is None or ln[1] + p * gs >= 3 * nroots:
                            ad.append(None)
                        else:
                            ad.append(ln[0] * scr + (ln[1] + p * gs) * tsz + col + 2 * ch)
                    wB += wavefronts128(ad)
                    idealB += sum(1 for q4 in range(4) if any(a is not None for a in ad[8 * q4:8 * q4 + 8]))
                if (lab + 1) % 2 == 1:
                    ad = []
                    for ln in lanes:
                        if ln is None or ln[1] + p * gs >= 3 * nroots:
                            ad.append(None)
                        else:
                            ad.append(ln[0] * scr + (ln[1] + p * gs) * tsz + col + lab)
                    wB += wavefronts64(ad)
                    idealB += sum(1 for h2 in range(2) if any(a is not None for a in ad[16 * h2:16 * h2 + 16]))
    # phase C: loads, per root, per m, per direction, per chunk
    nld = lab // 2 + 1
    for m in range(mk):
        for d in range(3):
            for ch in range(nld):
                ad = []
                for ln in lanes:
                    if ln is None:
                        ad.append(None)
                        continue
                    f = ln[1] * mk + m
                    if f >= ncd:
                        ad.append(None)
                        continue
                    ic, idd = f // s["nd"], f % s["nd"]
                    col = (cart_exps(lc, ic)[d] * (ld + 1) + cart_exps(ld, idd)[d]) * labp
                    ad.append(ln[0] * scr + d * tsz + col + 2 * ch)
                w = wavefronts128(ad)
                wC += w * nroots
                idealC += nroots * sum(1 for q4 in range(4) if any(a is not None for a in ad[8 * q4:8 * q4 + 8]))
    return wB, wC, idealB, idealC


def simulate_digestion(la, lb, lc, ld, mk, hs, scr, nk=1):
    """shared-memory wavefronts of the J/K digestion of one warp (64-bit accesses; depends on SCR only)"""
    s = sizes(la, lb, lc, ld, mk, hs)
    gs, qw, ncd, nc, nd, nap, nb, ne = s["gs"], s["qw"], s["ncd"], s["nc"], s["nd"], s["nap"], s["nb"], s["ne"]
    lanes = [(l // gs, l % gs) if l // gs < qw else None for l in range(32)]
    w = 0

    def acc(fn):
        nonlocal w
        w += wavefronts64([None if ln is None else fn(*ln) for ln in lanes])

    for e in range(ne):                                  # J(a,b) partials
        acc(lambda q, g: q * scr + e * gs + g)
    w += 2 * qw * gs * ((ne + 31) // 32)                 # reduction loads (lanes differ in e only)
    for _ in range(nk):
        for m in range(mk):
            for r in range(nap + nb):                    # half 1 stores
                acc(lambda q, g: None if g * mk + m >= ncd else q * scr + (r * nc + (g * mk + m) // nd) * nd + (g * mk + m) % nd)
        for p in range(((nap + nb) * nc + gs - 1) // gs):
            for l_ in range(nd):
                acc(lambda q, g: None if g + p * gs >= (nap + nb) * nc else q * scr + (g + p * gs) * nd + l_)
        for m in range(mk):
            for r in range(nap + nb):                    # half 2 stores
                acc(lambda q, g: None if g * mk + m >= ncd else q * scr + (r * nd + (g * mk + m) % nd) * nc + (g * mk + m) // nd)
        for p in range(((nap + nb) * nd + gs - 1) // gs):
            for k in range(nc):
                acc(lambda q, g: None if g + p * gs >= (nap + nb) * nd else q * scr + (g + p * gs) * nc + k)
    return w


DIG_WEIGHT = 0.3     # digestions per primitive quartet in the c18 warp-group classes (1.1 - 5 primitive quartets per shell quartet)


def constraints(la, lb, lc, ld, mk, hs, labp, tsz, scr):
    s = sizes(la, lb, lc, ld, mk, hs)
    lab = s["lab"]
    if labp % 2 or tsz % 2 or scr % 2:
        return False
    if labp < lab + 1 or (lab % 2 == 0 and labp < lab + 2):
        return False
    if tsz < (lc + 1) * (ld + 1) * labp:
        return False
    tq = s["nroots"] * 3 * tsz + s["rwp"]
    return scr >= max(tq, s["r1"], s["r2"])


def search(key, max_growth=1.12):
    la, lb, lc, ld, mk, hs, minb = shape(key)
    s = sizes(la, lb, lc, ld, mk, hs)
    cur = current_strides(la, lb, lc, ld, mk, hs)
    cb = simulate(la, lb, lc, ld, mk, hs, *cur)
    lab = s["lab"]
    labp0 = lab + 1 + ((lab + 1) % 2)
    if lab % 2 == 0:
        labp0 = lab + 2
    best = None
    for dl in range(0, 9, 2):
        labp = labp0 + dl
        t0 = (lc + 1) * (ld + 1) * labp
        for dt in range(0, 16, 2):
            tsz = t0 + dt
            tq = s["nroots"] * 3 * tsz + s["rwp"]
            s0 = max(tq, s["r1"], s["r2"])
            s0 += s0 % 2
            for ds in range(0, 16, 2):
                scr = s0 + ds
                if scr > cur[2] * max_growth + 8:
                    continue
                if not constraints(la, lb, lc, ld, mk, hs, labp, tsz, scr):
                    continue
                wB, wC, iB, iC = simulate(la, lb, lc, ld, mk, hs, labp, tsz, scr)
                wD = simulate_digestion(la, lb, lc, ld, mk, hs, scr)
                cand = (wB + wC + DIG_WEIGHT * wD, scr, labp, tsz, wB + wC, wD)
                if best is None or cand < best:
                    best = cand
    cb = cb + (simulate_digestion(la, lb, lc, ld, mk, hs, cur[2]),)
    return (la, lb, lc, ld, mk, hs, minb), cur, cb, best


def main():
    emit = "--emit" in sys.argv
    rows = []
    for key in sorted(CFG):
        shp, cur, cb, best = search(key)
        rows.append((key, shp, cur, cb, best))
    if not emit:
        print("class   H    S   MK GS | current LABP TSZ SCR : wavefronts B+C (ideal) | best LABP TSZ SCR : wavefronts | saved")
        for key, shp, cur, cb, best in rows:
            la, lb, lc, ld, mk, hs, minb = shp
            s = sizes(la, lb, lc, ld, mk, hs)
            print("%2d  %s%s|%s%s  MK %d GS %2d | %3d %4d %5d : %6d (%6d) dig %5d | %3d %4d %5d : %6d dig %5d | %4.0f%%" % (
                key % 100, NAMES[la], NAMES[lb], NAMES[lc], NAMES[ld], mk, s["gs"], cur[0], cur[1], cur[2], cb[0] + cb[1], cb[2] + cb[3], cb[4],
                best[2], best[3], best[1], best[4], best[5], 100.0 * (1 - best[0] / (cb[0] + cb[1] + DIG_WEIGHT * cb[4]))))
        return
    print("// generated by tools/wg_bank_model.py --emit : (LABP, TSZ, SCR) per warp-group configuration <LA,LB,LC,LD,MK,HS>")
    print("__host__ __device__ constexpr int wg_stride_key(int la, int lb, int lc, int ld, int mk, int hs) {")
    print("    return ((((la * 4 + lb) * 4 + lc) * 4 + ld) * 8 + mk) * 16 + hs;")
    print("}")
    print("// returns LABP | TSZ << 8 | SCR << 20, or 0 when the configuration is not in the table (formula strides are used)")
    print("__host__ __device__ constexpr long long wg_strides(int la, int lb, int lc, int ld, int mk, int hs) {")
    print("    switch (wg_stride_key(la, lb, lc, ld, mk, hs)) {")
    seen = set()
    for key, shp, cur, cb, best in rows:
        la, lb, lc, ld, mk, hs, minb = shp
        k = ((((la * 4 + lb) * 4 + lc) * 4 + ld) * 8 + mk) * 16 + hs
        if k in seen:
            continue
        seen.add(k)
        print("        case %d: return %dLL | (%dLL << 8) | (%dLL << 20);   // %s%s|%s%s MK %d HS %d: %d -> %d wavefronts" % (
            k, best[2], best[3], best[1], NAMES[la], NAMES[lb], NAMES[lc], NAMES[ld], mk, hs, cb[0] + cb[1], best[4]))
    print("        default: return 0;")
    print("    }")
    print("}")


if __name__ == "__main__":
    main()


def explore():
    """per-quartet wavefronts (A estimated with a random-row factor 2.6, B, C) for alternative MK / orientation"""
    import math
    for key in sorted(k for k in CFG if k < 100):
        bra, ket = CLS[key // 10], CLS[key % 10]
        out = []
        for swap in (0, 1):
            (la, lb), (lc, ld) = (ket, bra) if swap else (bra, ket)
            for mk in (1, 2, 3, 4, 6):
                ncd = ncart(lc) * ncart(ld)
                gs = (ncd + mk - 1) // mk
                if gs > 32 or mk > ncd:
                    continue
                for hs in (1, 2):
                    if ncart(la) % hs:
                        continue
                    s = sizes(la, lb, lc, ld, mk, hs)
                    if mk * s["ne"] > 72:
                        continue
                    cur = current_strides(la, lb, lc, ld, mk, hs)
                    CFG[-1] = (mk, 0, hs, 2)
                    best = None
                    # quick search
                    lab = s["lab"]
                    labp0 = lab + 2 if lab % 2 == 0 else lab + 1
                    for dl in (0, 2):
                        labp = labp0 + dl
                        t0 = (lc + 1) * (ld + 1) * labp
                        for dt in range(0, 16, 2):
                            tsz = t0 + dt
                            tq = s["nroots"] * 3 * tsz + s["rwp"]
                            s0 = max(tq, s["r1"], s["r2"]); s0 += s0 % 2
                            for ds in range(0, 16, 2):
                                wB, wC, iB, iC = simulate(la, lb, lc, ld, mk, hs, labp, tsz, s0 + ds)
                                c = (wB + wC, s0 + ds, labp, tsz)
                                if best is None or c < best:
                                    best = c
                    nv = 2 * s["nroots"]
                    passes = (nv + gs - 1) // gs
                    wA = 0
                    for p in range(passes):
                        act = [l for l in range(32) if l // gs < s["qw"] and (l % gs) + p * gs < nv]
                        wA += 7 * 2.6 * len(set(l // 8 for l in act))
                    tot = (best[0] + wA) * hs / s["qw"]
                    smem = best[1] * s["qw"] * 4 * 8
                    out.append((tot, swap, mk, hs, gs, s["qw"], mk * s["ne"], smem))
        out.sort()
        cur = CFG[key]
        print("%d %s%s|%s%s current (MK %d swap %d HS %d):" % (key, NAMES[bra[0]], NAMES[bra[1]], NAMES[ket[0]], NAMES[ket[1]], cur[0], cur[1], cur[2]),
              "  ".join("[%.0f wf/q sw%d MK%d HS%d GS%d QW%d acc%d %dK]%s" % (o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7] // 1024,
                                                                   "*" if (o[2], o[1], o[3]) == cur[:3] else "") for o in out[:6]))


if "--explore" in sys.argv:
    explore()
