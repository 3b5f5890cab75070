"""Python emulation of k_gemm2h's staged epilogue index math."""
import random
def swz(x): return x ^ ((((x >> 6) ^ (x >> 9) ^ (x >> 12)) & 7) << 3)
def run(tm, tn, lane_n_first, rng, force_fallback=False):
    nbr = tm + tn - 2
    tps_log = tm + tn - 6; s_log = 8 - tps_log; S = 1 << s_log
    rc = tm + tn + s_log + 2
    pos = list(range(rc)); rng.shuffle(pos)
    if rng.random() < 0.4:   # make the vector path likely: C bits 0,1,2 are tile bits
        pass
    c_shift = pos[:tm+tn]; grid_bits = pos[tm+tn:tm+tn+s_log]
    cbase = []
    for sb in range(S):
        v = 0
        for b in range(s_log):
            if (sb >> b) & 1: v |= 1 << grid_bits[b]
        cbase.append(v)
    if S > 1 and rng.random() < 0.5: cbase[S-1] = -1
    cs_mtop = c_shift[tm-1]; cs_ntop = c_shift[tm+tn-1]
    n0, n1 = tm - 1, tm
    def build(mp, mswap):
        # logical m positions: [first local bit, second local bit, the others ascending] (plan.cpp)
        order = ([mp, 0] if mswap else [0, mp]) + [q for q in range(1, tm-1) if q != mp]
        lm = {q: k for k, q in enumerate(order)}
        ent = sorted([(c_shift[i], lm[i]) for i in range(tm-1)] + [(c_shift[tm+i], (tm-1)+i) for i in range(tn-1)])
        e_cs = [e[0] for e in ent]; e_spos = [e[1] for e in ent]
        ecase = 0   # staging layout case: which tile bits are C bits 0,1,2
        if e_cs[:3] == [0, 1, 2] and e_spos[0] == 0:
            if e_spos[1:3] == [1, 2]: ecase = 1
            elif e_spos[1:3] == [1, n0]: ecase = 2
            elif e_spos[1:3] == [n0, 1]: ecase = 3
            elif e_spos[1:3] == [n0, n1]: ecase = 4
        if ecase >= 2:
            low = {2: (0, 1, 2, 3), 3: (0, 2, 1, 3), 4: (0, 3, 1, 2)}[ecase]
            def remap(sp):
                if sp == 0: return low[0]
                if sp == 1: return low[1]
                if sp == n0: return low[2]
                if sp == n1: return low[3]
                return sp + 2 if sp < n0 else sp
            e_spos = [remap(sp) for sp in e_spos]
        return ecase, e_cs, e_spos
    mp, mswap = 1, 0
    if tm >= 4:
        q_by_cs = {c_shift[q]: q for q in range(tm-1) if c_shift[q] < 3}
        q0 = q_by_cs.get(0, -1)
        if q0 > 0: mp, mswap = q0, 1
        elif q0 == 0:
            q1 = q_by_cs[1] if q_by_cs.get(1, -1) > 0 else q_by_cs.get(2, -1)
            if q1 > 1: mp = q1
    ecase, e_cs, e_spos = build(mp, mswap)
    if ecase == 0 and (mp != 1 or mswap):
        mp, mswap = 1, 0
        ecase, e_cs, e_spos = build(mp, mswap)
    evec = (e_cs[:3] == [0, 1, 2]) and not force_fallback
    CASES[(ecase, mp != 1, mswap)] = CASES.get((ecase, mp != 1, mswap), 0) + 1
    def m_phys(tmh, i, top):
        # the thread's accumulator row i (after the optional swap: bit 0 = first local bit, bit 1 = second) -> m
        tmw = ((tmh & ((1 << (mp-1)) - 1)) << 1) | ((tmh >> (mp-1)) << (mp+1))
        b0, b1 = (i & 1), (i >> 1) & 1
        if mswap: b0, b1 = b1, b0      # b0 = physical m0 bit, b1 = physical m_mp bit
        return tmw | b0 | (b1 << mp) | (top << (tm-1))
    out = {}
    stg = [[-1]*4096, [-1]*4096]
    def coords(ctid):
        sub = ctid >> tps_log; lt = ctid & ((1 << tps_log) - 1)
        if lane_n_first: tmh = lt >> (tn-3); tnh = lt & ((1 << (tn-3)) - 1)
        else: tmh = lt & ((1 << (tm-3)) - 1); tnh = lt >> (tm-3)
        return sub, tmh, tnh
    def val(sub, m, n): return (sub << 20) | (m << 10) | n
    for r in range(4):
        ih, jh = r & 1, r >> 1
        buf = stg[r & 1]
        for ctid in range(256):
            sub, tmh, tnh = coords(ctid)
            qbase = (sub << nbr) | (tmh << 2) | (tnh << (tm+1))
            qbase2 = (sub << nbr) | (tmh << 4) | (tnh << (tm+1))
            def V(i, j):
                mi = m_phys(tmh, i, ih)
                ni = (tnh*4 + j) if jh == 0 else ((1 << (tn-1)) + tnh*4 + j)
                return val(sub, mi, ni)
            W = [[(V(2*i1, j), V(2*i1+1, j)) for j in range(4)] for i1 in range(2)]   # packed words
            if ecase <= 1:
                for j in range(4):
                    a = swz(qbase | (j << (tm-1)))   # one 8-byte staging store
                    buf[a], buf[a+1], buf[a+2], buf[a+3] = W[0][j] + W[1][j]
            else:
                if ecase == 2:
                    v0 = [W[0][0], W[1][0], W[0][1], W[1][1]]; v1 = [W[0][2], W[1][2], W[0][3], W[1][3]]
                elif ecase == 3:
                    v0 = [W[0][0], W[0][1], W[1][0], W[1][1]]; v1 = [W[0][2], W[0][3], W[1][2], W[1][3]]
                else:
                    v0 = [W[0][0], W[0][1], W[0][2], W[0][3]]; v1 = [W[1][0], W[1][1], W[1][2], W[1][3]]
                for base, v in ((swz(qbase2), v0), (swz(qbase2 | 8), v1)):
                    for wi, wd in enumerate(v):
                        buf[base + 2*wi], buf[base + 2*wi + 1] = wd
        roff = (ih << cs_mtop) | (jh << cs_ntop)
        for ctid in range(256):
            if evec:
                ts = tc = 0
                for b in range(3, 11):
                    bit = (ctid >> (b-3)) & 1
                    if b < nbr: ts |= bit << e_spos[b]; tc |= bit << e_cs[b]
                for itr in range(2):
                    e8 = (itr << 11) | (ctid << 3)
                    sub_e = e8 >> nbr
                    so, co = ts, tc
                    if 11 < nbr: so |= itr << e_spos[11]; co |= itr << e_cs[11]
                    cb = cbase[sub_e]
                    if cb >= 0:
                        so |= sub_e << nbr
                        for q in range(8):
                            dq = ((q & 1) << e_spos[0]) | (((q >> 1) & 1) << e_spos[1]) | (((q >> 2) & 1) << e_spos[2])
                            a = cb + roff + co + q
                            assert a not in out, "double store"
                            if ecase != 0:   # one LDS.128
                                assert e_spos[:3] == [0, 1, 2]
                                out[a] = buf[swz(so) + q]
                            else:
                                out[a] = buf[swz(so | dq)]
            else:
                ts1 = tc1 = 0
                for b in range(0, 8):
                    bit = (ctid >> b) & 1
                    if b < nbr: ts1 |= bit << e_spos[b]; tc1 |= bit << e_cs[b]
                for itr in range(16):
                    e1 = (itr << 8) | ctid
                    sub_e = e1 >> nbr
                    so, co = ts1, tc1
                    for b in range(8, 12):
                        bit = (itr >> (b-8)) & 1
                        if b < nbr: so |= bit << e_spos[b]; co |= bit << e_cs[b]
                    cb = cbase[sub_e]
                    if cb >= 0:
                        a = cb + roff + co
                        assert a not in out, "double store"
                        out[a] = buf[swz(so | (sub_e << nbr))]
    n_expected = 0
    for sub in range(S):
        if cbase[sub] < 0: continue
        for m in range(1 << tm):
            for n in range(1 << tn):
                a = cbase[sub]
                for b in range(tm):
                    if (m >> b) & 1: a += 1 << c_shift[b]
                for b in range(tn):
                    if (n >> b) & 1: a += 1 << c_shift[tm+b]
                assert out.get(a) == val(sub, m, n), (tm, tn, sub, m, n, out.get(a))
                n_expected += 1
    assert len(out) == n_expected
CASES = {}
def run_case(tm, tn, lnf, rng, want):
    # steer the random C positions so that C bits 0,1,2 are the tile bits of layout case `want`
    class R(random.Random):
        pass
    hi = tm - 2    # highest in-round m bit (where the compiler pushes batch-in-child labels)
    nbr_bits = {1: [0, 1, 2], 2: [0, 1, tm], 3: [0, tm, 1], 4: [0, tm, tm + 1],
                5: [hi, tm, 0], 6: [hi, 0, tm], 7: [hi, tm, tm + 1], 8: [hi, 0, 1], 9: [0, hi, tm], 10: [0, tm, hi]}[want]   # indices into c_shift
    if want in (1,) and tm < 4: return
    if want >= 5 and tm < 5: return
    orig = rng.shuffle
    def shuffle(pos):
        orig(pos)
        for cbit, idx in enumerate(nbr_bits):   # swap so that c_shift[idx] == cbit
            k = pos.index(cbit); pos[k], pos[idx] = pos[idx], pos[k]
    rng.shuffle = shuffle
    try:
        run(tm, tn, lnf, rng)
    finally:
        del rng.shuffle
rng = random.Random(2)
for tm in range(3, 8):
    for tn in range(3, 8):
        if tm + tn < 9: continue
        for lnf in (0, 1):
            run(tm, tn, lnf, rng)
            run(tm, tn, lnf, rng, True)
            for want in range(1, 11):
                run_case(tm, tn, lnf, rng, want)
assert all(any(k[0] == c for k in CASES) for c in range(5)), CASES
assert any(k[1] for k in CASES) and any(k[2] for k in CASES), CASES
print("packed staged epilogue emulation OK")
