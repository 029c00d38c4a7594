"""Parity comparison rules shared by the GPU tests (BASELINE.json north_star):

  - integer / index work (keys, buckets, vote counts, argmax, cluster assignment) is bit-exact
    whenever it is fed the same floats;
  - float work may differ from the oracle by libm ulps; a quantised key may flip only when the
    feature lies within EDGE_TOL = 1e-5 of a bin edge (checked in double); the same tolerance
    covers the two other float *decisions* on the path: the source/target swap inside
    computePairFeatures and the d/2 radius predicate;
  - final poses agree within 1 mm / 0.5 degrees.
"""
import numpy as np

EDGE_TOL = 1e-5
FEAT_TOL = np.array([4e-6, 2e-6, 2e-6, 0.0])  # abs tolerance on f1..f4 (f4 is IEEE-exact)
ALPHA_TOL = 4e-6
POSE_T_TOL = 1e-3  # metres
POSE_R_TOL_DEG = 0.5


def circ_diff(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return np.minimum(d, 2 * np.pi - d)


def quantise(f, angle_step, dist_step):
    """PCL's floor(f/step) in float32, vectorised."""
    f = np.asarray(f, np.float32)
    a = np.float32(angle_step)
    d = np.float32(dist_step)
    q = np.empty(f.shape, np.int64)
    q[..., :3] = np.floor(f[..., :3] / a).astype(np.int64)
    q[..., 3] = np.floor(f[..., 3] / d).astype(np.int64)
    return q


def near_edge(f, angle_step, dist_step, tol=EDGE_TOL):
    """True where a feature component lies within tol of a quantisation edge (double)."""
    f = np.asarray(f, np.float64)
    steps = np.array([angle_step, angle_step, angle_step, dist_step], np.float64)
    r = f / steps
    return np.abs(r - np.round(r)) * steps < tol


def swap_margin(cloud, i, j):
    """|acos|a1| - acos|a2|| of computePairFeatures in double, for index arrays i, j."""
    p = cloud[:, :3].astype(np.float64)
    n = cloud[:, 3:6].astype(np.float64)
    d = p[j] - p[i]
    f4 = np.linalg.norm(d, axis=-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        a1 = np.abs(np.einsum("...k,...k->...", n[i], d) / f4)
        a2 = np.abs(np.einsum("...k,...k->...", n[j], d) / f4)
    return np.abs(np.arccos(np.clip(a1, -1, 1)) - np.arccos(np.clip(a2, -1, 1)))


def alpha_tolerance(cloud, i, j):
    """alpha = atan2(-z, y) of point j in the frame of (i): the frame's entries carry a few float
    ulps of sin/cos error (~|p| * 1e-7 absolute on y, z), which the angle magnifies by 1/rho, rho =
    distance of point j from the normal line through point i.  Tolerance = ALPHA_TOL + 6e-7*|p|/rho."""
    p = cloud[:, :3].astype(np.float64)
    n = cloud[:, 3:6].astype(np.float64)
    d = p[j] - p[i]
    along = np.einsum("...k,...k->...", d, n[i])
    rho = np.linalg.norm(d - along[..., None] * n[i], axis=-1)
    scale = np.maximum(np.linalg.norm(p[i], axis=-1), np.linalg.norm(p[j], axis=-1)) + 1e-3
    return ALPHA_TOL + 6e-7 * scale / np.maximum(rho, 1e-12)


def pose_error(A, B):
    """(translation distance [m], rotation angle [deg]) between two 4x4 / 3x4 poses."""
    A = np.asarray(A, np.float64).reshape(-1, 4)[:3]
    B = np.asarray(B, np.float64).reshape(-1, 4)[:3]
    dt = np.linalg.norm(A[:, 3] - B[:, 3])
    R = A[:, :3].T @ B[:, :3]
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    return dt, np.degrees(np.arccos(c))


def compare_features(cloud, F_dev, F_ref, angle_step, dist_step):
    """K1 rule.  Returns a dict of counters; raises AssertionError on a violation."""
    n = cloud.shape[0]
    F_dev = F_dev.reshape(n * n, 5)
    F_ref = F_ref.reshape(n * n, 5)
    nan_d = np.isnan(F_dev[:, 0])
    nan_r = np.isnan(F_ref[:, 0])
    assert np.array_equal(nan_d, nan_r), "validity masks (NaN rows) differ"
    v = np.flatnonzero(~nan_r)
    i, j = v // n, v % n
    # f4 is pure IEEE add/mul/sqrt: bit-exact
    assert np.array_equal(F_dev[v, 3], F_ref[v, 3]), "f4 must be bit-exact"
    ambiguous = swap_margin(cloud, i, j) < EDGE_TOL
    ok = ~ambiguous
    diff = np.abs(F_dev[v, :3].astype(np.float64) - F_ref[v, :3].astype(np.float64))
    diff[:, 0] = circ_diff(F_dev[v, 0], F_ref[v, 0])
    bad = (diff > FEAT_TOL[:3]).any(axis=1) & ok
    assert not bad.any(), f"{bad.sum()} pair features beyond tolerance, worst {diff[ok].max(axis=0)}"
    da = circ_diff(F_dev[v, 4], F_ref[v, 4])
    atol = alpha_tolerance(cloud, i, j)
    assert (da[ok] <= atol[ok]).all(), f"alpha_m beyond tolerance: {(da[ok] / atol[ok]).max()} x tol"
    # key agreement rule
    qd = quantise(F_dev[v, :4], angle_step, dist_step)
    qr = quantise(F_ref[v, :4], angle_step, dist_step)
    flip = (qd != qr)
    edge = near_edge(F_ref[v, :4], angle_step, dist_step)
    illegal = (flip & ~edge).any(axis=1) & ok
    assert not illegal.any(), f"{illegal.sum()} key flips away from a bin edge"
    return {"pairs": int(v.size), "swap_ambiguous": int(ambiguous.sum()), "key_flips": int(flip.any(axis=1).sum()),
            "max_feature_diff": diff[ok].max(axis=0).tolist(), "max_alpha_diff": float(da[ok].max())}


def table_buckets(table):
    """Device CSR -> dict {(d1,d2,d3,d4): (i array, j array, alpha array)} over all slices."""
    off, ei, ej, ea = table.export()
    ti = table.info
    ks = ti.key_space
    out = {}
    for s in range(ti.n_slices):
        o = off[s * ks:(s + 1) * ks + 1].astype(np.int64)
        ne = np.flatnonzero(o[1:] > o[:-1])
        keys = table.unpack_key(ne)
        for k, key in zip(ne, keys):
            sl = slice(o[k], o[k + 1])
            t = tuple(int(x) for x in key)
            if t in out:
                pi, pj, pa = out[t]
                out[t] = (np.concatenate([pi, ei[sl]]), np.concatenate([pj, ej[sl]]), np.concatenate([pa, ea[sl]]))
            else:
                out[t] = (ei[sl], ej[sl], ea[sl])
    return out
