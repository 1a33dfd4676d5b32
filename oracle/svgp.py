"""TEST INFRASTRUCTURE ONLY -- CPU oracle (NumPy/SciPy fp64) for the SVGP stage of ``GPDTW3D.fit``
(ensembles/models.py:357-424), SURVEY 8f rank 4.  Never imported by the product.

The arithmetic lives in GPflow 2.1.5 / TensorFlow 2.8.1 (un-vendored, not installable here); it is restated from the
published source (``gpflow/models/svgp.py``, ``gpflow/conditionals/util.py:base_conditional``,
``gpflow/kernels/stationaries.py``, ``gpflow/kullback_leiblers.py``, ``gpflow/optimizers/natgrad.py``) and anchored
on the reference's call site:

  models.py:358-364   kernel = Matern32(active_dims=[3]) + Matern32([0, 1]) + Matern32([2]) + Matern32([4 .. 4 + R))
  models.py:370       inducing_points = linspace(min(X, 0), max(X, 0), n_inducing)          (trainable, GPflow default)
  models.py:371-376   SVGP(kernel, _HeteroskedasticGaussian(), inducing_points, num_latent_gps=1)
                      -> whiten=True, q_mu = 0, q_sqrt = I, **num_data=None => the ELBO is NOT rescaled by N / batch**
  models.py:379-380   minibatches of ``minibatch_size`` from a shuffled, repeated dataset (unseeded in the reference)
  models.py:382-385   adam = Adam(0.01); natgrad = NaturalGradient(gamma=0.5); q_mu / q_sqrt not trainable by Adam
  models.py:388-391   each step: natgrad.minimize(loss, [(q_mu, q_sqrt)]) on ONE minibatch, then
                      adam.minimize(loss, trainable_variables) on the NEXT minibatch (the closure draws a new batch per call)
  models.py:393       n_optim_nits * (N // minibatch_size) steps
  models.py:408-411   mu, var = predict_f(X, full_cov=False); var += Y[:, 1]
  models.py:418-423   Distribution(mu, covariance=var, dist_type=dx.Normal)      (the variance goes in as a SCALE, Q-SCALE)

PARITY STATUS: unpinned against GPflow itself (the reference's minibatch order is unseeded, so even the reference does
not reproduce its own fit).  The MINIBATCH ORDER is defined here (``batch_indices``): a NumPy ``default_rng(seed)``
permutation of the N points per epoch, consumed ``minibatch_size`` at a time, epochs concatenated; the last,
incomplete slice of an epoch is carried into the next one.  The analytic gradient is checked against finite differences
of ``svgp_elbo`` in tests/test_oracle_svgp.py.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

SQRT3 = math.sqrt(3.0)
DEFAULT_JITTER = 1e-6
LOG_2PI = math.log(2.0 * math.pi)


def component_dims(n_realisations: int):
    """active_dims of the four Matern32 components in the order of models.py:358-364."""
    return [(3,), (0, 1), (2,), tuple(range(4, 4 + n_realisations))]


def softplus(u):
    return np.logaddexp(0.0, u)


def softplus_inv(x):
    return x + np.log(-np.expm1(-x))


def _r2(Xa, Xb, dims, ls):
    """GPflow ``square_distance`` on the active dimensions, scaled by the lengthscale (expansion form)."""
    A = Xa[:, dims] / ls
    B = Xb[:, dims] / ls
    return -2.0 * (A @ B.T) + np.sum(A * A, axis=1)[:, None] + np.sum(B * B, axis=1)[None, :]


def kernel_sum(Xa, Xb, variances, lengthscales, dims_list):
    """Sum of the four Matern-3/2 kernels: ``K = sum_c s2_c (1 + sqrt3 r_c) exp(-sqrt3 r_c)``, r clamped at 1e-18."""
    K = np.zeros((Xa.shape[0], Xb.shape[0]))
    for v, l, d in zip(variances, lengthscales, dims_list):
        r = np.sqrt(np.maximum(_r2(Xa, Xb, list(d), l), 1e-36))
        K += v * (1.0 + SQRT3 * r) * np.exp(-SQRT3 * r)
    return K


def conditional(Z, Xb, variances, lengthscales, dims_list, q_mu, q_sqrt, jitter=DEFAULT_JITTER):
    """``base_conditional(Kmn, Kmm, Knn_diag, f=q_mu, q_sqrt=q_sqrt, white=True, full_cov=False)``.
    Returns fmean [n], fvar [n] and (A = Lu^-1 Kuf [M, n], Lu)."""
    M = Z.shape[0]
    Kuu = kernel_sum(Z, Z, variances, lengthscales, dims_list) + jitter * np.eye(M)
    Lu = np.linalg.cholesky(Kuu)
    Kuf = kernel_sum(Z, Xb, variances, lengthscales, dims_list)
    A = sla.solve_triangular(Lu, Kuf, lower=True)
    kff = np.full(Xb.shape[0], float(np.sum(variances)))  # K(x, x) of a stationary sum kernel
    fmean = A.T @ q_mu
    W = np.tril(q_sqrt).T @ A
    fvar = kff - np.sum(A * A, axis=0) + np.sum(W * W, axis=0)
    return fmean, fvar, A, Lu


def svgp_elbo(Z, Xb, yb, sb, variances, lengthscales, dims_list, q_mu, q_sqrt, jitter=DEFAULT_JITTER):
    """``SVGP.elbo`` with ``num_data=None`` (scale 1) and the heteroskedastic Gaussian likelihood of models.py:142-149."""
    fmean, fvar, _, _ = conditional(Z, Xb, variances, lengthscales, dims_list, q_mu, q_sqrt, jitter)
    var_exp = -0.5 * LOG_2PI - 0.5 * np.log(sb) - 0.5 * ((yb - fmean) ** 2 + fvar) / sb
    M = Z.shape[0]
    d = np.diag(q_sqrt)
    kl = 0.5 * (np.sum(q_mu * q_mu) + np.sum(np.tril(q_sqrt) ** 2) - M - np.sum(np.log(d * d)))
    return float(np.sum(var_exp) - kl)


def natgrad_step(A, yb, sb, q_mu, q_sqrt, gamma=0.5):
    """``NaturalGradient(gamma).minimize(loss, [(q_mu, q_sqrt)])`` on one minibatch, XiNat parameterisation.
    With ``f_b = A^T v`` the expected log-likelihood is linear in the expectation parameters, so
    ``nat <- (1 - gamma) nat + gamma nat*``, ``nat1* = A D^-1 y``, ``nat2* = -1/2 (I + A D^-1 A^T)``."""
    M = A.shape[0]
    s_sqrt_inv = sla.solve_triangular(np.tril(q_sqrt), np.eye(M), lower=True)
    s_inv = s_sqrt_inv.T @ s_sqrt_inv
    nat1, nat2 = s_inv @ q_mu, -0.5 * s_inv
    Aw = A / sb[None, :]
    nat1_star = Aw @ yb
    nat2_star = -0.5 * (np.eye(M) + Aw @ A.T)
    nat1 = nat1 - gamma * (nat1 - nat1_star)
    nat2 = nat2 - gamma * (nat2 - nat2_star)
    nat2 = 0.5 * (nat2 + nat2.T)
    var_sqrt_inv = np.linalg.cholesky(-2.0 * nat2)
    var_sqrt = sla.solve_triangular(var_sqrt_inv, np.eye(M), lower=True)
    S = var_sqrt.T @ var_sqrt
    return S @ nat1, np.linalg.cholesky(S)


def _chol_backprop(L, Lbar):
    P = np.tril(L.T @ np.tril(Lbar))
    P[np.diag_indices_from(P)] *= 0.5
    S = sla.solve_triangular(L, P, lower=True, trans="T")
    S = sla.solve_triangular(L, S.T, lower=True, trans="T").T
    return 0.5 * (S + S.T)


def elbo_grads(Z, Xb, yb, sb, variances, lengthscales, dims_list, q_mu, q_sqrt, jitter=DEFAULT_JITTER):
    """d ELBO / d (variances [4], lengthscales [4], Z [M, d]) at fixed (q_mu, q_sqrt): what TensorFlow's tape gives
    ``adam.minimize(loss, trainable_variables)`` (models.py:391), up to the sign of the loss and the softplus chain.

    With m = A^T q_mu, v = kff - colsum(A o A) + colsum((Sq^T A) o (Sq^T A)):
      gm = (y - m) / s,  gv = -1 / (2 s),  Abar = q_mu gm^T + 2 (Sq Sq^T A - A) diag(gv),
      Kuf_bar = Lu^-T Abar,  Lu_bar = -tril(Lu^-T Abar A^T),  Kuu_bar = chol-backprop(Lu, Lu_bar),  kff_bar = gv.
    Matern-3/2 component c:  dK/ds2 = K_c / s2,  dK/dl = 3 s2 r^2 e^{-sqrt3 r} / l,
      dK(z, x)/dz_d = -3 s2 e^{-sqrt3 r} (z_d - x_d) / l^2  for d in the component's active dims (0 where r is clamped)."""
    M, dtot = Z.shape
    fmean, fvar, A, Lu = conditional(Z, Xb, variances, lengthscales, dims_list, q_mu, q_sqrt, jitter)
    gm = (yb - fmean) / sb
    gv = -0.5 / sb
    Sq = np.tril(q_sqrt)
    Abar = np.outer(q_mu, gm) + 2.0 * (Sq @ (Sq.T @ A) - A) * gv[None, :]
    Kuf_bar = sla.solve_triangular(Lu, Abar, lower=True, trans="T")
    Lu_bar = -np.tril(Kuf_bar @ A.T)
    Kuu_bar = _chol_backprop(Lu, Lu_bar)
    g_var = np.zeros(4)
    g_ls = np.zeros(4)
    g_Z = np.zeros_like(Z)
    for c, (v, l, d) in enumerate(zip(variances, lengthscales, dims_list)):
        d = list(d)
        for Kbar, Xo, sym in ((Kuf_bar, Xb, 1.0), (Kuu_bar, Z, 2.0)):
            r2 = _r2(Z, Xo, d, l)
            live = r2 > 1e-36
            r = np.sqrt(np.maximum(r2, 1e-36))
            e = np.exp(-SQRT3 * r)
            g_var[c] += np.sum(Kbar * (1.0 + SQRT3 * r) * e)
            g_ls[c] += np.sum(Kbar * np.where(live, 3.0 * v * r * r * e / l, 0.0))
            coef = np.where(live, -3.0 * v * e / (l * l), 0.0) * Kbar  # [M, n]
            # sum_n coef[m, n] (z_md - x_nd); Kuu is symmetric in its two arguments: twice the first-argument term
            g_Z[:, d] += sym * (coef.sum(axis=1)[:, None] * Z[:, d] - coef @ Xo[:, d])
        g_var[c] += np.sum(gv)  # kff = sum_c s2_c
    return g_var, g_ls, g_Z


class Adam:
    """tf.optimizers.Adam(0.01) of TF 2.8, elementwise over any array (models.py:382)."""

    def __init__(self, shape, lr=0.01, b1=0.9, b2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.m, self.v, self.t = np.zeros(shape), np.zeros(shape), 0

    def step(self, x, g):
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        lr_t = self.lr * math.sqrt(1 - self.b2**self.t) / (1 - self.b1**self.t)
        return x - lr_t * self.m / (np.sqrt(self.v) + self.eps)


def batch_indices(N, minibatch_size, n_batches, seed):
    """The DOCUMENTED minibatch order (module docstring): [n_batches, minibatch_size] int64."""
    rng = np.random.default_rng(seed)
    need = n_batches * minibatch_size
    stream = np.concatenate([rng.permutation(N) for _ in range(need // N + 2)])
    return stream[:need].reshape(n_batches, minibatch_size).astype(np.int64)


def inducing_points(X, n_inducing):
    """models.py:370."""
    return np.linspace(np.min(X, axis=0), np.max(X, axis=0), n_inducing)


def svgp_fit(X, Y, n_steps, n_inducing=400, minibatch_size=500, seed=0, gamma=0.5, lr=0.01, jitter=DEFAULT_JITTER,
             train_hypers=True, return_state=False):
    """``GPDTW3D.fit`` from the SVGP construction on (models.py:357-411): X [N, 4 + R], Y [N, 2] = (DTW mean, variance).
    ``n_steps`` = n_optim_nits * (N // minibatch_size) in the reference.  Returns (mu [N], var [N]) with var already
    including ``+ Y[:, 1]`` (models.py:411)."""
    N, dtot = X.shape
    R = dtot - 4
    dims_list = component_dims(R)
    Z = inducing_points(X, n_inducing)
    M = Z.shape[0]
    u_var = np.full(4, softplus_inv(1.0))
    u_ls = np.full(4, softplus_inv(1.0))
    q_mu, q_sqrt = np.zeros(M), np.eye(M)
    adam_var, adam_ls, adam_Z = Adam(4, lr), Adam(4, lr), Adam(Z.shape, lr)
    idx = batch_indices(N, minibatch_size, 2 * n_steps, seed)
    for step in range(n_steps):
        var, ls = softplus(u_var), softplus(u_ls)
        b = idx[2 * step]
        _, _, A, _ = conditional(Z, X[b], var, ls, dims_list, q_mu, q_sqrt, jitter)
        q_mu, q_sqrt = natgrad_step(A, Y[b, 0], Y[b, 1], q_mu, q_sqrt, gamma)  # models.py:390
        if train_hypers:
            b = idx[2 * step + 1]
            g_var, g_ls, g_Z = elbo_grads(Z, X[b], Y[b, 0], Y[b, 1], var, ls, dims_list, q_mu, q_sqrt, jitter)
            # loss = -ELBO; softplus chain: d softplus(u) / du = 1 - exp(-x)
            u_var = adam_var.step(u_var, -g_var * (-np.expm1(-var)))  # models.py:391
            u_ls = adam_ls.step(u_ls, -g_ls * (-np.expm1(-ls)))
            Z = adam_Z.step(Z, -g_Z)
    var, ls = softplus(u_var), softplus(u_ls)
    mu = np.empty(N)
    fv = np.empty(N)
    for c0 in range(0, N, 4096):
        m, v, _, _ = conditional(Z, X[c0:c0 + 4096], var, ls, dims_list, q_mu, q_sqrt, jitter)  # models.py:408
        mu[c0:c0 + 4096], fv[c0:c0 + 4096] = m, v
    fv = fv + Y[:, 1]  # models.py:411
    if return_state:
        return mu, fv, dict(variances=var, lengthscales=ls, Z=Z, q_mu=q_mu, q_sqrt=q_sqrt)
    return mu, fv
