"""TEST INFRASTRUCTURE: a CPU stand-in for ``Backend`` built on the oracle, so that the host-side
sharding / all-reduce logic of ``grid.fit_weight_barycentre_member_sharded`` can run under gloo on
a box without a GPU.  Never imported by the product."""
import types

import numpy as np
import torch

from oracle import reference_path as rp


class OracleOps:
    device = torch.device("cpu")

    def _in(self, t, shape=None, name="tensor"):
        t = torch.as_tensor(np.asarray(t), dtype=torch.float64).contiguous()
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), name
        return t

    def gpdtw1d_inputs(self, r):
        r = r.numpy()
        out = [rp.gpdtw1d_inputs(x) for x in r]
        return (torch.tensor(np.stack([o[0] for o in out])), torch.tensor(np.stack([o[1] for o in out])),
                torch.tensor(np.stack([o[2] for o in out])))

    def gp_posterior(self, X, ym, yv, var, ls, jitter=1e-6, want_cov=True, want_scale_tri=True):
        mus, vds, stats = [], [], []
        for b in range(X.shape[0]):
            mu, cov = rp.gp_posterior_closed_form(X[b].numpy(), ym[b].numpy(), yv[b].numpy(), float(var[b]), float(ls[b]),
                                                  jitter)
            L = np.linalg.cholesky(cov)
            import scipy.linalg as sla

            a = sla.solve_triangular(L, np.ones_like(mu), lower=True)
            bb = sla.solve_triangular(L, mu, lower=True)
            mus.append(mu)
            vds.append(np.diag(cov).copy())
            stats.append([a @ a, a @ bb, bb @ bb, np.log(np.diag(L)).sum()])
        B = X.shape[0]
        return types.SimpleNamespace(mu=torch.tensor(np.stack(mus)), var_diag=torch.tensor(np.stack(vds)),
                                     mvn_stats=torch.tensor(np.array(stats)),
                                     info_fit=torch.zeros(B, dtype=torch.int32), info_dist=torch.zeros(B, dtype=torch.int32))

    def loglik_weights_mvn(self, stats, obs, M, c=1.0, want_lls=False):
        C, Ro, T = obs.shape
        st = stats.numpy().reshape(C, M, 4)
        o = obs.numpy()
        ll = np.empty((C, M, Ro, T))
        for ci in range(C):
            for m in range(M):
                aa, ab, bb, ld = st[ci, m]
                ll[ci, m] = -0.5 * (o[ci] ** 2 * aa - 2 * o[ci] * ab + bb) - 0.5 * T * rp.LOG_2PI - ld
        lm = ll.mean(axis=2)
        with np.errstate(all="ignore"):
            le = np.exp(c * lm)
            w = le / np.nansum(le, axis=1, keepdims=True)
        w, le, lm = torch.tensor(w), torch.tensor(le), torch.tensor(lm)
        return (w, le, lm) if want_lls else w

    def barycentre_1d_partial(self, means, variances, lls_exp):
        return torch.stack([torch.nansum(lls_exp, 1), (lls_exp * means).sum(1), (lls_exp * variances.sqrt()).sum(1)])

    def weights_normalise(self, lls_exp, total):
        with np.errstate(all="ignore"):
            return lls_exp / total[:, None, :]

    def weights_time_mean(self, w):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = np.nanmean(w.numpy(), axis=2, keepdims=True)
        return torch.tensor(np.broadcast_to(m, w.shape).copy())

    def barycentre_1d_finish(self, partial, tolerance=1e-6, init_var=1.0, max_iters=200):
        s0, s1, s2 = partial.numpy()
        C, N = s0.shape
        mu, sd, it = np.empty((C, N)), np.empty((C, N)), np.empty((C, N), dtype=np.int32)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(all="ignore"):
                for c in range(C):
                    for i in range(N):
                        m_, s_, n_ = rp.gaussian_barycentre([s1[c, i] / s0[c, i]], [s2[c, i] / s0[c, i]], [1.0], tolerance,
                                                            init_var)
                        mu[c, i], sd[c, i], it[c, i] = m_, s_, n_
        return torch.tensor(mu), torch.tensor(sd), torch.tensor(it)
