"""Generates tests/golden/dba_reference.npz by EXECUTING the reference's own NumPy DBA code,
ensembles/dtwa.py (pure NumPy; imported by file path so that ensembles/__init__.py and its
uninstallable dependencies are not touched).  Run in the build container only:

    python tests/golden/make_golden_dba.py

The reference is read-only and never copied; only inputs and outputs are committed.
"""
import importlib.util
import os
import sys

import numpy as np

REF = os.environ.get("BE_REFERENCE", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_dtwa", os.path.join(REF, "ensembles", "dtwa.py"))
dtwa = importlib.util.module_from_spec(spec)
spec.loader.exec_module(dtwa)


def series_set(rng, R, T, kind):
    t = np.linspace(0.0, 1.0, T)
    if kind == "gmst":  # SURVEY 8d construction: trend + AR(1) noise per realisation
        g = rng.uniform(0.5, 4.0) * t + rng.uniform(0.0, 2.0) * t * t
        out = np.empty((R, T))
        for r in range(R):
            e = np.zeros(T)
            z = rng.normal(0.0, 0.12, T)
            for i in range(1, T):
                e[i] = 0.6 * e[i - 1] + z[i]
            out[r] = g + e
        return out
    if kind == "shifted":  # time-shifted bumps: the warping actually matters
        out = np.empty((R, T))
        for r in range(R):
            c = 0.3 + 0.4 * rng.random()
            out[r] = np.exp(-0.5 * ((t - c) / 0.08) ** 2) + 0.02 * rng.normal(size=T)
        return out
    if kind == "ties":  # small-integer values: exact ties in the DP exercise the tie rule
        return rng.integers(0, 3, size=(R, T)).astype(np.float64)
    raise ValueError(kind)


def main():
    rng = np.random.default_rng(20240 + 176)
    out = {}
    cases = [("gmst", 3, 40), ("gmst", 5, 57), ("shifted", 4, 48), ("shifted", 6, 33), ("ties", 4, 24), ("ties", 5, 31)]
    out["n_cases"] = np.array(len(cases))
    for c, (kind, R, T) in enumerate(cases):
        X = series_set(rng, R, T, kind)
        cost_mat = np.zeros((T, T))
        delta_mat = np.zeros((T, T))
        path_mat = np.zeros((T, T), dtype=np.int8)
        sq = np.array([[dtwa.squared_DTW(X[a], X[b], cost_mat, delta_mat) for b in range(R)] for a in range(R)])
        medoid = dtwa.approximate_medoid_index(list(X), cost_mat, delta_mat)
        one = dtwa.DBA_update(X[medoid], list(X), cost_mat, path_mat, delta_mat)
        centers = {n: dtwa.performDBA(list(X), n_iterations=n) for n in (1, 3, 10)}
        out[f"c{c}_X"] = X
        out[f"c{c}_kind"] = np.array(kind)
        out[f"c{c}_sqdtw"] = sq
        out[f"c{c}_medoid"] = np.array(medoid)
        out[f"c{c}_update1"] = one
        for n, v in centers.items():
            out[f"c{c}_center{n}"] = np.asarray(v)
        assert np.array_equal(one, centers[1])
        print(kind, R, T, "medoid", medoid, "sq range", sq[sq > 0].min() if (sq > 0).any() else 0.0, sq.max())
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dba_reference.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
