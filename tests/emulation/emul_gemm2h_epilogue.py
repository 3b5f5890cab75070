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
    ent = sorted([(c_shift[i], i) for i in range(tm-1)] + [(c_shift[tm+i], (tm-1)+i) for i in range(tn-1)])
    e_cs = [e[0] for e in ent]; e_spos = [e[1] for e in ent]
    cs_mtop = c_shift[tm-1]; cs_ntop = c_shift[tm+tn-1]
    evec = (e_cs[:3] == [0, 1, 2]) and not force_fallback
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
            for j in range(4):
                for i in range(4):       # 4 int16 = one 8-byte staging store
                    mi = (tmh*4 + i) if ih == 0 else ((1 << (tm-1)) + tmh*4 + i)
                    ni = (tnh*4 + j) if jh == 0 else ((1 << (tn-1)) + tnh*4 + j)
                    buf[swz(qbase | (j << (tm-1))) + i] = val(sub, mi, ni)
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
rng = random.Random(2)
for tm in range(3, 8):
    for tn in range(3, 8):
        if tm + tn < 9: continue
        for lnf in (0, 1):
            run(tm, tn, lnf, rng)
            run(tm, tn, lnf, rng, True)
print("packed staged epilogue emulation OK")
