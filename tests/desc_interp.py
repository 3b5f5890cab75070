"""Test-side interpreter of libtbcuda's raw step descriptors (csrc/desc.h).

It executes an exported plan with numpy exactly as the device kernels are specified to (same
address arithmetic: shift tables, panel bases, sentinel -2^30), so the HOST plan compiler --
label analysis, layouts, fused-subtree shared-memory stack, arena lifetimes, GEMM tiling tables --
can be checked against the oracle on a machine without a GPU.  Test infrastructure only.
"""
from __future__ import annotations

import numpy as np

NO_BIT = 0xFF
LOC_ARENA, LOC_POOL, LOC_SMEM = 0, 1, 2
KIND_GENERIC, KIND_GEMM = 1, 2
NEG_I32 = -(1 << 30)
NEG_I16 = -(1 << 14)
NEG_CFG = -(1 << 62)

SUBSTEP = np.dtype([("a_off", "<u2"), ("b_off", "<u2"), ("c_off", "<u2"), ("a_loc", "u1"), ("b_loc", "u1"),
                    ("c_loc", "u1"), ("rc", "u1"), ("nk", "u1"), ("nka", "u1"), ("nkb", "u1"), ("sa", "u1"),
                    ("sb", "u1"), ("pad", "u1"), ("a_shift", "u1", 16), ("b_shift", "u1", 16)])
SUBTREE = np.dtype([("out_off", "<i8"), ("first_step", "<u4"), ("n_steps", "<u4"), ("smem_elems", "<u4"),
                    ("pad", "<u4")])
BIGSTEP = np.dtype([("a_off", "<i8"), ("b_off", "<i8"), ("c_off", "<i8"), ("a_loc", "u1"), ("b_loc", "u1"),
                    ("kind", "u1"), ("rc", "u1"), ("nk", "u1"), ("nka", "u1"), ("nkb", "u1"), ("sa", "u1"),
                    ("sb", "u1"), ("tm", "u1"), ("tn", "u1"), ("kc", "u1"), ("ng", "u1"), ("n_mhi", "u1"),
                    ("n_nhi", "u1"), ("store_mode", "u1"), ("n_tiles", "<u4"), ("a_shift", "u1", 32),
                    ("b_shift", "u1", 32), ("c_shift", "u1", 32), ("ks", "u1"), ("po", "u1"), ("lane_n_first", "u1"), ("vec4", "u1")])
assert SUBSTEP.itemsize == 48 and SUBTREE.itemsize == 24 and BIGSTEP.itemsize == 144


def scatter(x: np.ndarray, shifts, n: int) -> np.ndarray:
    off = np.zeros_like(x)
    for i in range(n):
        s = int(shifts[i])
        if s != NO_BIT:
            off |= ((x >> i) & 1) << s
    return off


def _generic(A, B, rc, nk, nka, nkb, sa, sb, a_shift, b_shift, neg, kfirst=False):
    c = np.arange(1 << rc, dtype=np.int64)
    offA = scatter(c, a_shift, rc)
    offB = scatter(c, b_shift, rc)
    acc = np.full(1 << rc, neg, dtype=A.dtype)
    kmask = (1 << nk) - 1
    amask = (1 << (nk + nka)) - 1
    for r in range(1 << (nk + nka + nkb)):
        if kfirst:  # K0-first operand layouts: reduction bit 0 = address bit 0, the rest starts at sa+1 / sb+1
            assert nka == 0 and nkb == 0
            ra = (r & 1) | ((r >> 1) << (sa + 1))
            rb = (r & 1) | ((r >> 1) << (sb + 1))
            acc = np.maximum(acc, A[offA + ra] + B[offB + rb])
            continue
        ra = (r & amask) << sa
        rb = ((r & kmask) | ((r >> (nk + nka)) << nk)) << sb
        acc = np.maximum(acc, A[offA + ra] + B[offB + rb])
    return acc


def _rank_a(s):
    # rank of A = labels of C present in A + reduced labels living in A
    if s["kind"] == KIND_GEMM:
        return int(s["tm"]) + int(s["nk"]) + int(s["n_mhi"]) + (int(s["ng"]) - int(s["n_mhi"]) - int(s["n_nhi"]))
    return sum(1 for i in range(int(s["rc"])) if s["a_shift"][i] != NO_BIT) + int(s["nk"]) + int(s["nka"])


def _rank_b(s):
    if s["kind"] == KIND_GEMM:
        return int(s["tn"]) + int(s["nk"]) + int(s["n_nhi"]) + (int(s["ng"]) - int(s["n_mhi"]) - int(s["n_nhi"]))
    return sum(1 for i in range(int(s["rc"])) if s["b_shift"][i] != NO_BIT) + int(s["nk"]) + int(s["nkb"])


def run_plan(plan):
    """plan: tbcuda.Plan.  Returns (root value as float, arena array, dict of stats)."""
    hdr = np.frombuffer(plan.raw(5), dtype="<i8")
    arena_elems, root_off, n_levels, vt = (int(x) for x in hdr)
    dt = np.int64 if vt in (1, 3, 5) else np.float64 if vt == 4 else np.float32
    neg = NEG_I32 if vt == 1 else NEG_I16 if vt == 3 else NEG_CFG if vt == 5 else -np.inf
    if vt == 3:
        pool = np.frombuffer(plan.raw(0), dtype="<i2").astype(np.int64)
    elif vt == 4:   # Tropical{Float64}
        pool = np.frombuffer(plan.raw(0), dtype="<f8").copy()
    elif vt == 5:   # size << 32 | vertex mask
        pool = np.frombuffer(plan.raw(0), dtype="<i8").copy()
    else:
        praw = np.frombuffer(plan.raw(0), dtype="<u4")
        pool = praw.view("<i4").astype(np.int64) if vt == 1 else praw.view("<f4").copy()
    mt = 3  # log2 of a thread's microtile extent (8 x 8 outputs)
    sub = np.frombuffer(plan.raw(1), dtype=SUBSTEP)
    trees = np.frombuffer(plan.raw(2), dtype=SUBTREE)
    big = np.frombuffer(plan.raw(3), dtype=BIGSTEP)
    lvl = np.frombuffer(plan.raw(4), dtype="<i4")
    arena = np.full(max(arena_elems, 1), 12345 if vt in (1, 3, 5) else np.nan, dtype=dt)

    with np.errstate(invalid="ignore"):
        for t in trees:
            smem = np.full(max(int(t["smem_elems"]), 1), 777 if vt in (1, 3, 5) else np.nan, dtype=dt)
            for s in sub[int(t["first_step"]): int(t["first_step"]) + int(t["n_steps"])]:
                A = (smem if s["a_loc"] == LOC_SMEM else pool)[int(s["a_off"]):]
                B = (smem if s["b_loc"] == LOC_SMEM else pool)[int(s["b_off"]):]
                rc = int(s["rc"])
                acc = _generic(A, B, rc, int(s["nk"]), int(s["nka"]), int(s["nkb"]), int(s["sa"]), int(s["sb"]),
                               s["a_shift"], s["b_shift"], neg, kfirst=bool(s["pad"]))
                if s["c_loc"] == LOC_SMEM:
                    o = int(s["c_off"])
                    assert o + (1 << rc) <= int(t["smem_elems"]), "fused step writes past its subtree's shared memory"
                    smem[o:o + (1 << rc)] = acc
                else:
                    o = int(t["out_off"])
                    arena[o:o + (1 << rc)] = acc
        for lv in range(1, n_levels + 1):
            pending = []
            reads, writes = [], []
            for s in big[lvl[lv]:lvl[lv + 1]]:
                if s["a_loc"] == LOC_ARENA:
                    reads.append((int(s["a_off"]), int(s["a_off"]) + (1 << _rank_a(s))))
                if s["b_loc"] == LOC_ARENA:
                    reads.append((int(s["b_off"]), int(s["b_off"]) + (1 << _rank_b(s))))
                writes.append((int(s["c_off"]), int(s["c_off"]) + (1 << int(s["rc"]))))
            for w0, w1 in writes:
                for r0, r1 in reads:
                    assert w1 <= r0 or r1 <= w0, "a step's output overlaps an operand read in the same level"
            ws = sorted(writes)
            for (a0, a1), (b0, b1) in zip(ws, ws[1:]):
                assert a1 <= b0, "two outputs of one level overlap"
            for s in big[lvl[lv]:lvl[lv + 1]]:
                A = (pool if s["a_loc"] == LOC_POOL else arena)[int(s["a_off"]):]
                B = (pool if s["b_loc"] == LOC_POOL else arena)[int(s["b_off"]):]
                rc, nk = int(s["rc"]), int(s["nk"])
                if s["kind"] == KIND_GENERIC:
                    acc = _generic(A, B, rc, nk, int(s["nka"]), int(s["nkb"]), int(s["sa"]), int(s["sb"]),
                                   s["a_shift"], s["b_shift"], neg, kfirst=bool(s["store_mode"]))
                    if s["vec4"]:
                        assert int(s["po"]) in (10, 12) and int(s["ks"]) == 0 and rc >= int(s["po"])
                    else:
                        assert int(s["po"]) + int(s["ks"]) <= 8
                    assert int(s["po"]) <= rc and int(s["n_tiles"]) == 1 << (rc - int(s["po"]))
                    assert int(s["ks"]) <= nk + int(s["nka"]) + int(s["nkb"])
                    pending.append((int(s["c_off"]), np.arange(1 << rc), acc))
                else:
                    tm, tn, ng = int(s["tm"]), int(s["tn"]), int(s["ng"])
                    n_mhi, n_nhi = int(s["n_mhi"]), int(s["n_nhi"])
                    assert int(s["nka"]) == 0 and int(s["nkb"]) == 0 and tm + tn + ng == rc
                    moff = scatter(np.arange(1 << tm, dtype=np.int64), s["c_shift"], tm)
                    noff = scatter(np.arange(1 << tn, dtype=np.int64), s["c_shift"][tm:], tn)
                    S = 1 << (8 - (tm - mt + tn - 3))
                    assert 1 <= S <= 32 and tm >= mt and tn >= 3 and tm <= 7 and tn <= 7
                    assert int(s["n_tiles"]) == ((1 << ng) + S - 1) // S
                    assert S * (1 << int(s["kc"])) * ((1 << tm) + (1 << tn)) <= 4096 * (2 if vt == 3 else 1) and int(s["kc"]) <= nk
                    if s["store_mode"] == 1:
                        assert s["c_shift"][0] == 0 and s["c_shift"][1] == 1
                    if s["store_mode"] == 2:
                        assert s["c_shift"][tm] == 0 and s["c_shift"][tm + 1] == 1
                    # staged-epilogue tables: sorted in-round tile bits
                    nbr = tm + tn - 2
                    cs = [int(x) for x in s["c_shift"]]
                    n0, n1 = tm - 1, tm

                    def build(mp, mswap):
                        # logical m positions: [first local bit, second local bit, the others ascending]
                        order = ([mp, 0] if mswap else [0, mp]) + [q for q in range(1, tm - 1) if q != mp]
                        lm = {q: k for k, q in enumerate(order)}
                        ent = sorted([(cs[i], lm[i]) for i in range(tm - 1)] + [(cs[tm + i], (tm - 1) + i) for i in range(tn - 1)])
                        sp = [e[1] for e in ent]
                        ecase = 0
                        if vt == 3:
                            if [e[0] for e in ent[:3]] == [0, 1, 2] and sp[0] == 0:
                                ecase = {(1, 2): 1, (1, n0): 2, (n0, 1): 3, (n0, n1): 4}.get((sp[1], sp[2]), 0)
                            if ecase >= 2:
                                low = {2: (0, 1, 2, 3), 3: (0, 2, 1, 3), 4: (0, 3, 1, 2)}[ecase]
                                m = {0: low[0], 1: low[1], n0: low[2], n1: low[3]}
                                sp = [m[x] if x in m else (x + 2 if x < n0 else x) for x in sp]
                                assert sorted(sp) == list(range(nbr)) and sp[:3] == [0, 1, 2]
                        else:
                            ecase = int(sp[:2] == [0, 1])
                        return ecase, ent, sp

                    mp, mswap = 1, 0
                    if vt == 3 and tm >= 4:
                        q_by_cs = {cs[q]: q for q in range(tm - 1) if cs[q] < 3}
                        q0 = q_by_cs.get(0, -1)
                        if q0 > 0:
                            mp, mswap = q0, 1
                        elif q0 == 0:
                            q1 = q_by_cs[1] if q_by_cs.get(1, -1) > 0 else q_by_cs.get(2, -1)
                            if q1 > 1:
                                mp = q1
                    ecase, ent, sp = build(mp, mswap)
                    if ecase == 0 and (mp != 1 or mswap):
                        mp, mswap = 1, 0
                        ecase, ent, sp = build(mp, mswap)
                    assert (int(s["a_shift"][29]), int(s["b_shift"][29])) == (mp, mswap)
                    assert [int(x) for x in s["b_shift"][:nbr]] == [e[0] for e in ent]
                    assert int(s["a_shift"][30]) == cs[tm - 1] and int(s["a_shift"][31]) == cs[tm + tn - 1]
                    assert int(s["b_shift"][30]) == ecase
                    assert [int(x) for x in s["a_shift"][:nbr]] == sp
                    if vt == 3:
                        assert int(s["b_shift"][31]) == int([e[0] for e in ent[:3]] == [0, 1, 2])
                    else:
                        assert int(s["b_shift"][31]) == int(ent[0][0] == 0 and ent[1][0] == 1)
                    idxs, vals = [], []
                    for g in range(1 << ng):
                        gm = g & ((1 << n_mhi) - 1)
                        gn = (g >> n_mhi) & ((1 << n_nhi) - 1)
                        gb = g >> (n_mhi + n_nhi)
                        ab = (gm | (gb << n_mhi)) << (tm + nk)
                        bb = (gn | (gb << n_nhi)) << (tn + nk)
                        cb = int(scatter(np.array([g], dtype=np.int64), s["c_shift"][tm + tn:], ng)[0])
                        if vt == 3:  # packed int16: panels are [k_rest][tile index][k0]
                            assert int(s["kc"]) >= 1
                            Ap = A[ab:ab + (1 << (tm + nk))].reshape(1 << (nk - 1), 1 << tm, 2).transpose(0, 2, 1).reshape(1 << nk, 1 << tm)
                            Bp = B[bb:bb + (1 << (tn + nk))].reshape(1 << (nk - 1), 1 << tn, 2).transpose(0, 2, 1).reshape(1 << nk, 1 << tn)
                        else:
                            Ap = A[ab:ab + (1 << (tm + nk))].reshape(1 << nk, 1 << tm)
                            Bp = B[bb:bb + (1 << (tn + nk))].reshape(1 << nk, 1 << tn)
                        Ct = np.full((1 << tm, 1 << tn), neg, dtype=dt)
                        for k in range(1 << nk):
                            Ct = np.maximum(Ct, Ap[k][:, None] + Bp[k][None, :])
                        idxs.append((cb + moff[:, None] + noff[None, :]).reshape(-1))
                        vals.append(Ct.reshape(-1))
                    idx = np.concatenate(idxs)
                    assert len(np.unique(idx)) == (1 << rc) and idx.max() < (1 << rc), "gemm store map is not a bijection"
                    pending.append((int(s["c_off"]), idx, np.concatenate(vals)))
            # all steps of a level read before any writes land (they run concurrently on the device)
            for off, idx, vals in pending:
                arena[off + idx] = vals
    root = arena[root_off]
    if vt == 3:
        rootf = -np.inf if root <= -(1 << 13) else float(root)
    elif vt == 1:
        rootf = -np.inf if root <= -(1 << 29) else float(root)
    elif vt == 5:
        rootf = -np.inf if (int(root) >> 32) <= -(1 << 29) else float(int(root) >> 32)
    else:
        rootf = float(root)
    return rootf, arena


def to_float(arr, vt):
    if vt == 5:  # the size half of size + configuration elements
        size = arr >> 32
        out = size.astype(np.float64)
        out[size <= -(1 << 29)] = -np.inf
        return out
    if vt in (1, 3):
        out = arr.astype(np.float64)
        out[arr <= (-(1 << 29) if vt == 1 else -(1 << 13))] = -np.inf
        return out
    return arr.astype(np.float64)


def to_config(arr):
    """the vertex mask half of size + configuration elements (0 where the element is tropical zero)"""
    cfg = (arr & 0xFFFFFFFF).astype(np.uint32)
    cfg[(arr >> 32) <= -(1 << 29)] = 0
    return cfg
