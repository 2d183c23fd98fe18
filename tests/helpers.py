"""Shared helpers for the test-suite."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def spec_for(name):
    from oracle import problems
    return problems.demo_spec() if name == "demo" else problems.spec(name)


def rel_err(a, b):
    """max |a-b| relative to the inf-norm of the reference vector b (SURVEY 8d parity metric, per vector)"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return (np.max(np.abs(a - b)) if a.size else 0.0) / scale


def seg_rel_err(flat, a, b, dual):
    """worst per-segment relative error between two compact vectors (segments of DESIGN.md 'Data layout')"""
    n, m, nx, nu, L = flat.n, flat.m, flat.nx, flat.nu, flat.nleaf
    if dual:
        sizes = [flat.ysz, m, (n - 1) * nx, (n - 1) * nu, n - 1, n - 1, m * (nx + nu) if flat.nl_rect else 0,
                 L * nx, L, L, L * nx if flat.leaf_rect else 0]
    else:
        sizes = [n * nx, m * nu, flat.ysz, n, n]
    cuts = np.concatenate(([0], np.cumsum(sizes)))
    assert cuts[-1] == a.size == b.size
    worst = 0.0
    for k in range(len(sizes)):
        sa, sb = a[cuts[k]: cuts[k + 1]], b[cuts[k]: cuts[k + 1]]
        if sa.size == 0:
            continue
        ref = np.max(np.abs(sb))
        err = np.max(np.abs(sa - sb))
        worst = max(worst, err / ref if ref > 1e-12 else err)
    return worst
