"""One cfg2-shaped posterior call, for ncu launch lists (developer tool)."""
import sys, torch
sys.path.insert(0, '.')
from bayesian_ensembling_b200.backend import Backend
from bayesian_ensembling_b200 import synthetic as syn
be = Backend.get()
cfg = syn.CONFIGS['cfg2']
M, T = cfg.members, int(sys.argv[1]) if len(sys.argv) > 1 else cfg.steps
g = torch.Generator(device='cuda').manual_seed(0)
tn = torch.linspace(0, 1, T, device='cuda', dtype=torch.float64)
r = (2 * tn + tn**2)[None, None, :] + 0.12 * torch.randn(M, cfg.realisations, T, device='cuda', dtype=torch.float64, generator=g)
X, ym, yv = be.gpdtw1d_inputs(r)
var = torch.full((M,), 0.5, dtype=torch.float64, device='cuda'); ls = torch.full((M,), 6.0, dtype=torch.float64, device='cuda')
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for _ in range(reps):
    post = be.gp_posterior(X, ym, yv, var, ls, want_cov=False, want_scale_tri=False)
torch.cuda.synchronize()
print('ok', post.info_fit.sum().item(), post.info_dist.sum().item())
