"""One gp_posterior call on B problems of size T, for ncu captures (developer tool).
usage: prof_posterior.py [T=3012] [B=24] [reps=1]"""
import sys, torch
sys.path.insert(0, '.')
from bayesian_ensembling_b200.backend import Backend
be = Backend.get()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 3012
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
g = torch.Generator(device='cuda').manual_seed(0)
tn = torch.linspace(0, 1, T, device='cuda', dtype=torch.float64)
r = (2 * tn + tn**2)[None, None, :] + 0.12 * torch.randn(B, 5, T, device='cuda', dtype=torch.float64, generator=g)
X, ym, yv = be.gpdtw1d_inputs(r)
var = torch.full((B,), 0.5, dtype=torch.float64, device='cuda'); ls = torch.full((B,), 6.0, dtype=torch.float64, device='cuda')
for _ in range(reps):
    post = be.gp_posterior(X, ym, yv, var, ls, want_cov=False, want_scale_tri=False)
torch.cuda.synchronize()
print('ok', post.info_fit.sum().item(), post.info_dist.sum().item())
