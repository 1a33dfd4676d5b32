"""Developer parity/timing probe run on the GPU box (not a test; tests/ holds the real ones)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from bayesian_ensembling_b200.backend import Backend
from bayesian_ensembling_b200 import synthetic as syn
from oracle import reference_path as rp

be = Backend.get()
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

def check(T, M=3, R=4, Ro=3, seed=0):
    cfg = syn.Config("t", 9, 1, M, R, T, Ro, False, "")
    reals, obs = syn.make_cells(cfg, seed=seed)
    r = torch.tensor(reals[0], device='cuda')
    X, ym, yv = be.gpdtw1d_inputs(r)
    var = torch.full((M,), 0.5, dtype=torch.float64, device='cuda'); ls = torch.full((M,), 6.0, dtype=torch.float64, device='cuda')
    K = be.matern32_gram(X, var, ls)
    post = be.gp_posterior(X, ym, yv, var, ls)
    w, le, lm = be.loglik_weights_mvn(post.mvn_stats, torch.tensor(obs, device='cuda'), M, want_lls=True)
    bmu, bsd, bit = be.barycentre_1d(post.mu[None], post.var_diag[None], w)
    torch.cuda.synchronize()
    o = rp.cell_pipeline_L1(reals[0], obs[0], 0.5, 6.0)
    Xo, yo, so = rp.gpdtw1d_inputs(reals[0][0])
    res = dict(T=T,
        X=rel(X[0].cpu().numpy(), Xo), yv=rel(yv[0].cpu().numpy(), so),
        K=rel(K[0].cpu().numpy(), rp.matern32_gram(Xo, 0.5, 6.0)),
        mu=rel(post.mu.cpu().numpy(), o['mu']), cov=rel(post.cov.cpu().numpy(), o['cov']),
        var=float(np.abs(post.var_diag.cpu().numpy()/np.array([np.diag(c) for c in o['cov']])-1).max()),
        tri=rel(post.scale_tri.cpu().numpy(), o['scale_tri']),
        lls=float(np.abs(lm[0].cpu().numpy()-o['lls_mean']).max()/np.abs(o['lls_mean']).max()),
        info=(post.info_fit.tolist(), post.info_dist.tolist()))
    wn = w[0].cpu().numpy(); ok = ~np.isnan(o['weights'])
    res['w'] = float(np.abs(wn[ok]-o['weights'][ok]).max()) if ok.any() else None; res['w_nan_match'] = bool((np.isnan(wn) == np.isnan(o['weights'])).all())
    res['bmu'] = float(np.nanmax(np.abs(bmu[0].cpu().numpy()-o['bary_mu']))) if ok.any() else None
    res['bsd'] = float(np.nanmax(np.abs(bsd[0].cpu().numpy()-o['bary_std']))) if ok.any() else None
    print(json.dumps(res))

for T in [24, 86, 126, 128, 165, 251, 300, 600]:
    check(T)

# timing at cfg2 shape
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
cfg = syn.CONFIGS['cfg2']
reals, obs = syn.make_cells(cfg)
r = torch.tensor(reals[0], device='cuda'); ob = torch.tensor(obs, device='cuda')
M, T = cfg.members, cfg.steps
X, ym, yv = be.gpdtw1d_inputs(r)
var = torch.full((M,), 0.5, dtype=torch.float64, device='cuda'); ls = torch.full((M,), 6.0, dtype=torch.float64, device='cuda')
ms = timeit(lambda: be.gp_posterior(X, ym, yv, var, ls, want_cov=False, want_scale_tri=False))
print('cfg2 gp_posterior ms', ms, 'TF/s (4/3 T^3)', M * 4/3 * T**3 / ms / 1e9)
K = be.matern32_gram(X, var, ls); A = K + torch.diag_embed(yv) 
ms = timeit(lambda: be.potrf(A)); print('cfg2 potrf ms', ms, 'TF/s', M * T**3 / 3 / ms / 1e9)
ms = timeit(lambda: be.matern32_gram(X, var, ls)); print('cfg2 gram ms', ms)
post = be.gp_posterior(X, ym, yv, var, ls, want_cov=False, want_scale_tri=False)
def tail():
    w = be.loglik_weights_mvn(post.mvn_stats, ob, M); return be.barycentre_1d(post.mu[None], post.var_diag[None], w)
print('cfg2 weights+bary ms', timeit(tail))
print('info', post.info_fit.tolist(), post.info_dist.tolist())
