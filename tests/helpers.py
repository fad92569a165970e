"""Shared helpers for the parity tests (oracle = checker, engine = product path through the C-ABI)."""
import numpy as np

import oracle
from bundle_adjustment_solver_b200 import solver as S


def load_oracle(sc):
    return S.load_scene(oracle.FullBAOracle(), sc)


def load_engine(sc, device=0, identical_internal=None):
    """identical_internal: (T12, X) taken from the oracle so both sides start from bit-identical
    internal parameters (removes the 1-ulp freedom of the host-side pose inversion)."""
    e = S.load_scene(S.FullBundleAdjustmentSolver(device=device), sc)
    e._upload(internal_override=identical_internal)
    return e


def blockwise_rel_err(got, ref, block):
    """max over blocks of ||got-ref||_F / ||ref||_F, with blocks whose reference norm is below
    1e-13 of the largest block norm compared absolutely against that scale."""
    got = np.asarray(got).reshape(-1, block)
    ref = np.asarray(ref).reshape(-1, block)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    nr = np.linalg.norm(ref, axis=1)
    scale = nr.max() if nr.size else 0.0
    den = np.maximum(nr, 1e-13 * scale) + 1e-300
    return float((np.linalg.norm(got - ref, axis=1) / den).max())


def options_pair(**kw):
    """Same options for oracle and engine."""
    from bundle_adjustment_solver_b200.capi import default_options
    eo = default_options(**kw)
    okw = {k: v for k, v in kw.items() if k not in ("inverse_scaler", "check_every", "use_graph")}
    oo = oracle.default_full_options(**okw)
    return oo, eo


def S_block_view(Sflat, N):
    """(n*n,) column-major symmetric -> (N*N, 36) array of 6x6 blocks (row-major inside a block)."""
    n = 6 * N
    Sm = np.asarray(Sflat).reshape(n, n).T  # symmetric anyway
    return Sm.reshape(N, 6, N, 6).transpose(0, 2, 1, 3).reshape(N * N, 36)
