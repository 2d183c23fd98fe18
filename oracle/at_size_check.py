"""Compare compact iterates of a CUDA run with an at-size fixture recorded from the unmodified reference
(oracle/make_golden_at_size.py).  TEST INFRASTRUCTURE: a checker for tests/ and for bench.py's parity_check leg."""
import numpy as np


def check_against_fixture(flat, g, sfx, k, p, d):
    """p, d: compact GPU iterates of one instance after k iterations; g: fixture; sfx: '' or '_i<instance>'"""
    mp = flat.maps()
    # ---- sampled entries -----------------------------------------------------------------------------------------------
    pidx, didx = g["pidx" + sfx], g["didx" + sfx]
    pe, de = g["p_edges" + sfx], g["d_edges" + sfx]
    ps, ds = g[f"ps{k}{sfx}"], g[f"ds{k}{sfx}"]
    assert pe[-1] == flat.np_ and de[-1] == mp["d_total"]
    seg_of_p = np.searchsorted(pe, pidx, side="right") - 1
    err = np.abs(p[pidx] - g[f"p{k}{sfx}"]) / np.maximum(ps[seg_of_p, 0], 1e-300)
    worst = float(err.max())
    inv = np.full(mp["d_total"], -1, dtype=np.int64)
    inv[mp["d_gather"]] = np.arange(flat.nd_)
    real = inv[didx] >= 0
    want = g[f"d{k}{sfx}"]
    assert np.all(want[~real] == 0.0), "reference placeholders moved"
    seg_of_d = np.searchsorted(de, didx, side="right") - 1
    err = np.abs(d[inv[didx[real]]] - want[real]) / np.maximum(ds[seg_of_d[real], 0], 1e-300)
    worst = max(worst, float(err.max()))
    # ---- checksums of every segment -----------------------------------------------------------------------------------
    n, m, nx, nu, L = flat.n, flat.m, flat.nx, flat.nu, flat.nleaf
    psz = [n * nx, m * nu, flat.ysz, n, n]
    dsz = [flat.ysz, m, (n - 1) * nx, (n - 1) * nu, n - 1, n - 1, m * (nx + nu) if flat.nl_rect else 0,
           L * nx, L, L, L * nx if flat.leaf_rect else 0]
    for vec, sizes, stats in ((p, psz, ps), (d, dsz, ds)):
        cuts = np.concatenate(([0], np.cumsum(sizes)))
        assert cuts[-1] == vec.size
        for s in range(len(sizes)):
            seg = vec[cuts[s]: cuts[s + 1]]
            if seg.size == 0:
                continue
            amax, l1, l2, tot = stats[s]
            if amax < 1e-300:   # the reference segment is exactly zero (e.g. the rectangle duals while nothing is active):
                # ours may carry rounding dust of the Moreau step alpha (w - clip(w)); bound it against the vector's scale
                assert np.max(np.abs(seg)) <= 1e-9 * max(1e-300, float(np.max(stats[:, 0])))
                continue
            worst = max(worst, abs(np.max(np.abs(seg)) - amax) / amax, abs(np.sum(np.abs(seg)) - l1) / l1,
                        abs(np.sqrt(np.sum(seg * seg)) - l2) / l2, abs(np.sum(seg) - tot) / l1)
    return worst
