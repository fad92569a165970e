"""Partition plan of the banded reduced solve (csrc/ba_nd_plan.h), checked on the CPU: the fronts are emulated in
numpy exactly as csrc/ba_cholesky_nd.cuh processes them (assembly through the per-child index maps, partial
Cholesky of the own block, contribution block, backward substitution root first) and the result is compared with
numpy.linalg.solve.  Replaces Am_BCinvBt_mat.ldlt().solve(...) (core/full_bundle_adjustment_solver.cpp:890-908)."""
import ctypes as C

import numpy as np
import pytest

from bundle_adjustment_solver_b200 import capi

F = ["own0", "k", "k8", "rb0", "wr", "lb0", "wl", "b8", "child0", "child1", "parent", "rb_off", "lb_off",
     "rhs_off", "level", "cta", "seq", "L_off", "U_off", "helper"]


def nd_plan(N, b, max_ctas=128, depth=-1, chunk=-1):
    L = capi.lib()
    meta = np.zeros(8, dtype=np.int64)
    cap = 4096
    nodes = np.zeros((cap, 20), dtype=np.int64)
    nn = L.ba_debug_nd_plan(N, b, max_ctas, depth, chunk, capi.ptr(meta), capi.ptr(nodes), cap)
    assert nn <= cap
    return dict(valid=bool(meta[0]), depth=int(meta[1]), n_leaves=int(meta[2]), n_levels=int(meta[3]),
                n_ctas=int(meta[4]), max_tiles=int(meta[5]), max_R8=int(meta[6]), smem=int(meta[7])), \
        [dict(zip(F, map(int, row))) for row in nodes[:nn]]


def banded_spd(N, b, rng):
    """Block-banded SPD matrix: pose j couples with poses j-b .. j+b."""
    n = 6 * N
    G = rng.normal(size=(n, n))
    A = np.zeros((n, n))
    for j in range(N):
        lo = max(0, j - b)
        A[6 * j:6 * j + 6, 6 * lo:6 * j + 6] = G[6 * j:6 * j + 6, 6 * lo:6 * j + 6]
    A = np.tril(A)
    A = A + A.T
    A += np.eye(n) * (np.abs(A).sum(axis=1).max() + 1.0)
    return A


def emulate(nodes, S, rhs):
    n = len(rhs)
    U = {}
    Lf = {}
    order = sorted(range(len(nodes)), key=lambda i: nodes[i]["level"])

    def glob(nd, i):
        k8 = nd["k8"]
        if i < nd["k"]:
            return nd["own0"] + i
        bi = i - k8
        if bi < 0:
            return None
        if bi < nd["wr"]:
            return nd["rb0"] + bi
        if bi < nd["wr"] + nd["wl"]:
            return nd["lb0"] + bi - nd["wr"]
        if bi == nd["wr"] + nd["wl"]:
            return n
        return None

    Saug = np.zeros((n + 1, n + 1))
    Saug[:n, :n] = S
    Saug[n, :n] = rhs
    Saug[:n, n] = rhs
    for t in order:
        nd = nodes[t]
        R8 = nd["k8"] + nd["b8"]
        Fm = np.zeros((R8, R8))
        for j in range(nd["k8"]):
            if j >= nd["k"]:
                Fm[j, j] = 1.0
                continue
            gj = nd["own0"] + j
            for i in range(j, R8):
                gi = glob(nd, i)
                if gi is not None:
                    Fm[i, j] = Saug[gi, gj]
        for c in ("child0", "child1"):
            ci = nd[c]
            if ci < 0:
                continue
            cn = nodes[ci]
            assert cn["parent"] == t
            pm = {}
            for i in range(cn["wr"]):
                pm[cn["rb_off"] + i] = i
            for i in range(cn["wl"]):
                pm[cn["lb_off"] + i] = cn["wr"] + i
            pm[cn["rhs_off"]] = cn["wr"] + cn["wl"]
            Uc = U[ci]
            for i, pi in pm.items():
                for j, pj in pm.items():
                    if i >= j:
                        Fm[i, j] += Uc[max(pi, pj), min(pi, pj)]
        Fs = np.tril(Fm) + np.tril(Fm, -1).T
        k8 = nd["k8"]
        L11 = np.linalg.cholesky(Fs[:k8, :k8])
        L21 = np.linalg.solve(L11, Fs[:k8, k8:]).T
        Lf[t] = (L11, L21)
        U[t] = Fs[k8:, k8:] - L21 @ L21.T
    x = np.zeros(n)
    for t in reversed(order):
        nd = nodes[t]
        R8 = nd["k8"] + nd["b8"]
        xs = np.zeros(R8)
        for i in range(nd["k8"], R8):
            gi = glob(nd, i)
            if gi is not None:
                xs[i] = -1.0 if gi == n else x[gi]
        L11, L21 = Lf[t]
        tb = L21.T @ xs[nd["k8"]:]
        xo = -np.linalg.solve(L11.T, tb)
        x[nd["own0"]:nd["own0"] + nd["k"]] = xo[:nd["k"]]
    return x


@pytest.mark.parametrize("N,b,depth,chunk", [(198, 11, -1, -1), (64, 3, 2, -1), (90, 2, 3, 4), (61, 5, 1, 7),
                                             (400, 5, -1, -1), (47, 1, 2, 3), (150, 13, -1, -1)])
def test_partition_plan_solves_banded_system(N, b, depth, chunk):
    rng = np.random.default_rng(N * 31 + b)
    meta, nodes = nd_plan(N, b, depth=depth, chunk=chunk)
    assert meta["valid"], meta
    assert meta["smem"] <= 227 * 1024
    # every column is owned exactly once
    owned = np.zeros(6 * N, dtype=int)
    for nd in nodes:
        owned[nd["own0"]:nd["own0"] + nd["k"]] += 1
    assert (owned == 1).all()
    S = banded_spd(N, b, rng)
    rhs = rng.normal(size=6 * N)
    x = emulate(nodes, S, rhs)
    xr = np.linalg.solve(S, rhs)
    assert np.abs(x - xr).max() / np.abs(xr).max() < 1e-10
    # persistent driver: every node belongs to exactly one CTA list; inside a list a front follows its child0 (the
    # CTA climbs the tree), a front whose children belong to other CTAs starts a list; all CTAs are co-resident
    assert meta["n_ctas"] <= 148
    helpers = sorted(nd["helper"] for nd in nodes if nd["helper"] >= 0)
    assert helpers == list(range(meta["n_ctas"], meta["n_ctas"] + len(helpers))) and meta["n_ctas"] + len(helpers) <= 148
    assert all(nd["b8"] >= 96 for nd in nodes if nd["helper"] >= 0)          # only large boundary blocks get a helper
    lists = {}
    for t, nd in enumerate(nodes):
        lists.setdefault(nd["cta"], []).append((nd["seq"], t))
    assert sorted(lists) == list(range(meta["n_ctas"]))
    for c, lst in lists.items():
        lst.sort()
        assert [q for q, _ in lst] == list(range(len(lst)))
        for (_, a), (_, bnode) in zip(lst, lst[1:]):
            assert nodes[bnode]["child0"] == a
        first = nodes[lst[0][1]]
        for ch in ("child0", "child1"):
            assert first[ch] < 0 or nodes[first[ch]]["cta"] != c
    for t, nd in enumerate(nodes):
        for c in ("child0", "child1"):
            if nd[c] >= 0:
                assert nodes[nd[c]]["level"] < nd["level"]
        if nd["child1"] >= 0:
            assert nodes[nd["child1"]]["cta"] != nd["cta"]


def test_partition_plan_rejects_wide_bands_and_short_chains():
    assert not nd_plan(200, 15)[0]["valid"]      # boundary accumulators would not fit in registers
    assert not nd_plan(20, 8)[0]["valid"]        # two leaves of >= b poses do not fit
    meta, _ = nd_plan(1998, 5)
    assert meta["valid"] and meta["n_ctas"] <= 144
