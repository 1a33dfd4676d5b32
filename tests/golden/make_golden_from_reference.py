"""Generates tests/golden/prefit_members.npz from the reference's own pickled fits.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden_from_reference.py

Source: /root/reference/experiments/pre_fit_models/*.pkl, written by
``ModelCollection.save`` (ensembles/data.py:397-404) after
``ModelCollection.fit(GPDTW1D(), n_optim_nits=2500)``
(experiments/pre_fitting_cmip6models.py:76-77).  Per fitted member we keep the
INPUT realisations [R,T] and the reference's OUTPUTS mu [T], covariance [T,T]
(lower triangle, packed; the stored matrices are symmetric to the bit or to
~1e-18, the asymmetry is recorded) and distrax's Cholesky factor _scale_tri
(lower triangle, packed).  Also stores the kernel hyper-parameters recovered by
a least-squares fit of the closed-form posterior to the stored covariance
(structural pin of the GP stage).
"""
import glob
import os
import sys

import numpy as np
import scipy.optimize as so

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.fixtures import load_fitted_collection  # noqa: E402
from oracle import reference_path as rp  # noqa: E402

REF = "/root/reference/experiments/pre_fit_models"


def recover_hypers(reals, cov):
    X, y, s = rp.gpdtw1d_inputs(reals)
    target = cov - np.diag(s)

    def resid(p):
        var, ls = np.exp(p)
        _, c = rp.gp_posterior_closed_form(X, y, s, var, ls)
        return (c - np.diag(s) - target).ravel()

    best = None
    for p0 in ([np.log(0.3), np.log(5.0)], [np.log(1.0), np.log(2.0)], [np.log(0.1), np.log(8.0)]):
        r = so.least_squares(resid, p0, xtol=1e-14, ftol=1e-14, gtol=1e-14)
        if best is None or r.cost < best.cost:
            best = r
    var, ls = np.exp(best.x)
    return var, ls, np.abs(resid(best.x)).max()


def main():
    out = {}
    names = []
    for path in sorted(glob.glob(os.path.join(REF, "*.pkl"))):
        tag = os.path.basename(path).replace("_1D_models.pkl", "")
        for i, m in enumerate(load_fitted_collection(path)):
            key = f"{tag}.{i}"
            names.append(f"{key}|{m.name}")
            T = m.mu.shape[0]
            il = np.tril_indices(T)
            out[f"{key}.realisations"] = m.realisations
            out[f"{key}.mu"] = m.mu
            out[f"{key}.cov_lower"] = m.covariance[il]
            out[f"{key}.cov_asym"] = np.array(np.abs(m.covariance - m.covariance.T).max())
            out[f"{key}.scale_tri_lower"] = m.scale_tri[il]
            assert np.abs(np.triu(m.scale_tri, 1)).max() == 0.0
            var, ls, res = recover_hypers(m.realisations, m.covariance)
            out[f"{key}.recovered_hypers"] = np.array([var, ls, res])
            print(key, m.name, m.realisations.shape, "asym", out[f"{key}.cov_asym"], "hypers", var, ls, "resid", res)
    out["names"] = np.array(names)
    dst = os.path.join(os.path.dirname(__file__), "prefit_members.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) / 1e6, "MB")


if __name__ == "__main__":
    main()
