"""Generates tests/golden/wasserstein_reference.npz by EXECUTING the reference's own function
bodies for SURVEY rows a5-a8 (and the Stein kernel of f-3) in the build container:

    python tests/golden/make_golden_wasserstein.py

``ensembles/wasserstein.py`` imports tensorflow / distrax / jax at module level, none of which is
installable here, but the functions on the path are short pure-``jnp`` bodies.  Their SOURCE TEXT is
therefore cut out of the reference files with ``ast`` and ``exec``-ed under ``jnp = numpy`` plus a few
stand-ins for what the bodies touch (a distribution with ``mean() / covariance() / variance()``; for
``Barycentre._compute`` a model collection, ``xr.ones_like`` and a recording ``Distribution``; for
the KSD kernel ``jax.vmap`` / ``jax.jit`` / ``lax.scan`` as plain loops).  Nothing is copied into the
repository: only seeded inputs and the outputs the reference's code produced are committed.

  wasserstein.py:10-13    sqrtm                           -> a7
  wasserstein.py:15-19    wasserstien_distance            -> a8 (covariance term alone)
  wasserstein.py:21-47    gaussian_w2_distance_distrax    -> a8
  wasserstein.py:61-100   gaussian_barycentre             -> a5
  ensemble_scheme.py:43-81  Barycentre._compute           -> a6
  weights.py:360-393      k_0_fun / imq_KSD (KSDWeight)   -> f-3
"""
import ast
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("BE_REFERENCE", "/root/reference")


def _function_source(path, name, inside=None):
    """Source text of function ``name`` (optionally nested inside class / function ``inside``)."""
    text = open(path).read()
    tree = ast.parse(text)
    scope = tree
    for outer in inside or ():
        scope = next(n for n in ast.walk(scope) if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name == outer)
    node = next(n for n in ast.walk(scope) if isinstance(n, ast.FunctionDef) and n.name == name)
    node.decorator_list = []
    import textwrap

    seg = ast.get_source_segment(text, node)
    return textwrap.dedent(" " * node.col_offset + seg), (node.lineno, node.end_lineno)


class _JnpShim:
    """``jax.numpy`` names the extracted bodies use, on NumPy fp64 (the reference enables x64)."""

    ndarray = np.ndarray
    DeviceArray = np.ndarray

    def __getattr__(self, name):
        return getattr(np, name)


class _Dist:
    """What gaussian_w2_distance_distrax asks of a distrax distribution."""

    def __init__(self, mu, cov):
        self._mu, self._cov = np.asarray(mu, dtype=np.float64), np.asarray(cov, dtype=np.float64)

    def mean(self):
        return self._mu

    def covariance(self):
        return self._cov

    def variance(self):
        return np.diag(self._cov).copy()


def load_reference_functions():
    jnp = _JnpShim()
    ns = {"jnp": jnp, "np": np, "warnings": warnings, "distrax": type("distrax", (), {"Distribution": object}),
          "tfd": type("tfd", (), {"Distribution": object})}
    lines = {}
    wpath = os.path.join(REF, "ensembles", "wasserstein.py")
    for name in ("sqrtm", "wasserstien_distance", "gaussian_w2_distance_distrax", "gaussian_barycentre"):
        src, span = _function_source(wpath, name)
        exec(compile(src, f"{wpath}:{name}", "exec"), ns)
        lines[name] = span

    # ---- Barycentre._compute (ensemble_scheme.py:43-81) under stand-ins ------------------------------
    class _Recorded:
        def __init__(self, mu, covariance, dim_array, dist_type):
            self.mu, self.covariance, self.dim_array, self.dist_type = mu, covariance, dim_array, dist_type

    class _Blank:
        def __mul__(self, other):
            return self

        def rename(self, name):
            return self

    xr = type("xr", (), {"ones_like": staticmethod(lambda a: _Blank()), "DataArray": object})
    dx = type("dx", (), {"MultivariateNormalDiag": "MultivariateNormalDiag"})
    ns_b = dict(ns)
    ns_b.update({"xr": xr, "dx": dx, "Distribution": _Recorded, "trange": range, "ModelCollection": object,
                 "abc": __import__("abc")})
    spath = os.path.join(REF, "ensembles", "ensemble_scheme.py")
    src, span = _function_source(spath, "_compute", inside=("Barycentre",))
    exec(compile(src, f"{spath}:Barycentre._compute", "exec"), ns_b)
    lines["Barycentre._compute"] = span
    ns["barycentre_compute"] = ns_b["_compute"]

    # ---- the Stein kernel of KSDWeight (weights.py:360-393) --------------------------------------------
    def vmap(f, in_axes):
        def g(*args):
            n = next(np.asarray(a).shape[0] for a, ax in zip(args, in_axes) if ax == 0)
            return np.asarray([f(*[(np.asarray(a)[i] if ax == 0 else a) for a, ax in zip(args, in_axes)])
                               for i in range(n)])
        return g

    def scan(body, init, xs):
        carry = init
        for i in range(np.asarray(xs[0]).shape[0]):
            carry, _ = body(carry, tuple(np.asarray(x)[i] for x in xs))
        return carry, None

    jax = type("jax", (), {"vmap": staticmethod(vmap), "jit": staticmethod(lambda f: f)})
    lax = type("lax", (), {"scan": staticmethod(scan)})
    ns_k = dict(ns)
    ns_k.update({"jax": jax, "lax": lax})
    kpath = os.path.join(REF, "ensembles", "weights.py")
    src, span = _function_source(kpath, "k_0_fun", inside=("KSDWeight", "_compute"))
    exec(compile(src, f"{kpath}:k_0_fun", "exec"), ns_k)
    lines["k_0_fun"] = span
    ns_k["_batch_k_0_fun_rows"] = vmap(ns_k["k_0_fun"], (None, 0, None, 0, None, None))  # weights.py:377
    src, span2 = _function_source(kpath, "imq_KSD", inside=("KSDWeight", "_compute"))
    exec(compile(src, f"{kpath}:imq_KSD", "exec"), ns_k)
    lines["imq_KSD"] = span2
    ns["imq_KSD"] = ns_k["imq_KSD"]
    return ns, lines


def _spd(rng, T, scale=1.0, floor=1e-3):
    A = rng.normal(size=(T, T + 3))
    return scale * (A @ A.T / T + floor * np.eye(T))


def _posterior_like(rng, T):
    """degC-anomaly-like smooth covariance + heteroskedastic diagonal (what GPDTW1D posteriors look like)."""
    t = np.arange(T)[:, None]
    r = np.abs(t - t.T) / rng.uniform(3.0, 9.0)
    K = rng.uniform(2e-3, 8e-3) * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)
    return K + np.diag(rng.uniform(5e-3, 3e-2, size=T)), rng.normal(0.5, 0.4, size=T)


class _Model:
    def __init__(self, mu, var, R, T):
        class _D:
            pass

        d = _D()
        d._dist = type("dist", (), {"mean": staticmethod(lambda: mu), "variance": staticmethod(lambda: var)})

        class _MD:
            size = R * T
            realisation = type("r", (), {"size": R})

            def __getitem__(self, i):
                return type("row", (), {"drop": staticmethod(lambda name: None)})

        self.distribution, self.model_data = d, _MD()


class _Collection(list):
    @property
    def number_of_models(self):
        return len(self)


def main():
    fn, lines = load_reference_functions()
    rng = np.random.default_rng(20240 + 61)
    out = {"lines": np.array(repr(lines))}

    # a7: sqrtm
    sq_cases = [("spd", 1), ("spd", 2), ("spd", 5), ("spd", 24), ("spd", 86), ("posterior", 40), ("posterior", 165),
                ("singular", 12)]
    out["sqrtm_n"] = np.array(len(sq_cases))
    for i, (kind, T) in enumerate(sq_cases):
        if kind == "spd":
            A = _spd(rng, T, scale=rng.uniform(0.1, 10.0))
        elif kind == "posterior":
            A, _ = _posterior_like(rng, T)
        else:  # rank-deficient PSD
            Bm = rng.normal(size=(T, T // 2))
            A = Bm @ Bm.T
        out[f"sqrtm{i}_A"] = A
        out[f"sqrtm{i}_root"] = fn["sqrtm"](A)
        out[f"sqrtm{i}_kind"] = np.array(kind)

    # a8: W2 "distance", full and diagonal covariance, and the covariance-only variant
    w2_cases = [1, 3, 24, 86, 165]
    out["w2_n"] = np.array(len(w2_cases))
    for i, T in enumerate(w2_cases):
        if T >= 24:
            S1, m1 = _posterior_like(rng, T)
            S2, m2 = _posterior_like(rng, T)
        else:
            S1, S2 = _spd(rng, T), _spd(rng, T, scale=2.0)
            m1, m2 = rng.normal(size=T), rng.normal(size=T)
        a, b = _Dist(m1, S1), _Dist(m2, S2)
        out[f"w2_{i}_mu1"], out[f"w2_{i}_S1"], out[f"w2_{i}_mu2"], out[f"w2_{i}_S2"] = m1, S1, m2, S2
        out[f"w2_{i}_full"] = np.array(fn["gaussian_w2_distance_distrax"](a, b, full_cov=True))
        out[f"w2_{i}_diag"] = np.array(fn["gaussian_w2_distance_distrax"](a, b, full_cov=False))
        out[f"w2_{i}_self"] = np.array(fn["gaussian_w2_distance_distrax"](a, a, full_cov=True))
        out[f"w2_{i}_covonly"] = np.array(fn["wasserstien_distance"](S1, S2))

    # a5: gaussian_barycentre in every regime of the signed stop rule
    bc = []
    M = 6
    for regime in ("anomaly", "anomaly", "climb", "climb", "slow", "huge", "nan_weight", "tol", "init"):
        means = rng.normal(0.5, 0.4, size=M)
        w = rng.random(M)
        w /= w.sum()
        kw = {}
        if regime == "anomaly":      # sum w s < 1: exits at iteration 0 with variance = sum w s
            sd = rng.uniform(0.05, 0.2, size=M)
        elif regime == "climb":      # sum w s > 1: climbs towards (sum w s)^2
            sd = rng.uniform(1.5, 4.0, size=M)
        elif regime == "slow":       # barely above 1: many iterations
            sd = np.full(M, 1.003)
        elif regime == "huge":      # (sum w s)^2 ~ 1e11: still converges (the log-error halves per iteration)
            sd = rng.uniform(2e5, 4e5, size=M)
        elif regime == "nan_weight":
            sd = rng.uniform(0.05, 0.2, size=M)
            w = np.full(M, np.nan)
        elif regime == "tol":
            sd = rng.uniform(1.5, 4.0, size=M)
            kw = {"tolerance": 1e-3}
        else:
            sd = rng.uniform(0.5, 2.0, size=M)
            kw = {"init_var": 0.01}
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            _stdout = sys.stdout
            sys.stdout = open(os.devnull, "w")  # the reference prints on non-convergence (wasserstein.py:96)
            try:
                mu, sigma = fn["gaussian_barycentre"](means, sd, w, **kw)
            finally:
                sys.stdout.close()
                sys.stdout = _stdout
        bc.append((regime, means, sd, w, kw.get("tolerance", 1e-6), kw.get("init_var", 1.0), float(mu), float(sigma),
                   len(caught) > 0))
    out["bary_n"] = np.array(len(bc))
    for i, (regime, means, sd, w, tol, iv, mu, sigma, warned) in enumerate(bc):
        out[f"bary{i}_regime"] = np.array(regime)
        out[f"bary{i}_means"], out[f"bary{i}_sd"], out[f"bary{i}_w"] = means, sd, w
        out[f"bary{i}_tol"], out[f"bary{i}_init"] = np.array(tol), np.array(iv)
        out[f"bary{i}_mu"], out[f"bary{i}_sigma"], out[f"bary{i}_warned"] = np.array(mu), np.array(sigma), np.array(warned)
        print("barycentre", regime, mu, sigma, "warned" if warned else "")

    # a6: Barycentre._compute over M members x T points (variances mix the < 1 and > 1 regimes)
    M, T, R = 5, 16, 3
    mus = rng.normal(0.5, 0.4, size=(M, T))
    var = np.where(rng.random((M, T)) < 0.7, rng.uniform(0.005, 0.05, size=(M, T)), rng.uniform(2.0, 9.0, size=(M, T)))
    w = rng.random((M, T))
    w /= w.sum(0)
    coll = _Collection(_Model(mus[m], var[m], R, T) for m in range(M))
    weights = type("w", (), {"values": w})
    dist = fn["barycentre_compute"](None, coll, weights)
    assert dist.dist_type == "MultivariateNormalDiag"
    out["scheme_mus"], out["scheme_var"], out["scheme_w"] = mus, var, w
    out["scheme_mu_out"], out["scheme_covariance_out"] = np.asarray(dist.mu, dtype=np.float64), np.asarray(dist.covariance, dtype=np.float64)

    # f-3: the IMQ kernel Stein discrepancy of KSDWeight, samples [N,1] and score of Normal(mean, scale = variance)
    ksd_cases = [(3, 0.4, 0.02), (10, 0.7, 0.015), (10, -0.2, 0.5), (25, 1.0, 2.0)]
    out["ksd_n"] = np.array(len(ksd_cases))
    for i, (N, mean, scale) in enumerate(ksd_cases):
        samples = rng.normal(mean + 0.1, 0.15, size=(N, 1))
        grads = -(samples - mean) / (scale * scale)  # d/dx log N(x | mean, scale) (weights.py:421-423: scale = variance)
        out[f"ksd{i}_samples"], out[f"ksd{i}_mean"], out[f"ksd{i}_scale"] = samples, np.array(mean), np.array(scale)
        out[f"ksd{i}_value"] = np.array(fn["imq_KSD"](samples, grads))
        print("ksd", N, float(out[f"ksd{i}_value"]))

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wasserstein_reference.npz")
    np.savez_compressed(dst, **out)
    print("reference lines executed:", lines)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
