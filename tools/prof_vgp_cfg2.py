"""Two iterations of the L2 training loop (be_vgp_fit) on one cfg2 cell (24 members, T = 3012), for an ncu launch list."""
import sys

import torch

sys.path.insert(0, ".")
from bayesian_ensembling_b200 import synthetic  # noqa: E402
from bayesian_ensembling_b200.backend import Backend  # noqa: E402

be = Backend.get()
cfg = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
reals, _ = synthetic.make_cells(cfg, n_cells=1)
r = torch.as_tensor(reals, device=be.device)
C, M, R, T = r.shape
X, ym, yv = be.gpdtw1d_inputs(r.reshape(C * M, R, T))
post, var, ls = be.vgp_fit(X, ym, yv, 2, want_scale_tri=False)
torch.cuda.synchronize()
print("ok", int(post.info_fit.max()))
