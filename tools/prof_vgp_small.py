"""The L2 training loop (be_vgp_fit) at the small-T shapes (cfg1 / cfg4: T = 251; the reference's own fits: T = 165),
B member problems at a time: ms per iteration, and (under ncu) the launch list of a few iterations (developer tool)."""
import sys

import torch

sys.path.insert(0, ".")
from bayesian_ensembling_b200 import synthetic  # noqa: E402
from bayesian_ensembling_b200.backend import Backend  # noqa: E402

be = Backend.get()
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = synthetic.CONFIGS["cfg4"]
reals, _ = synthetic.make_cells(cfg, n_cells=cells)
r = torch.as_tensor(reals, device=be.device)
C, M, R, T = r.shape
X, ym, yv = be.gpdtw1d_inputs(r.reshape(C * M, R, T))
be.vgp_fit(X, ym, yv, 1, want_scale_tri=False)
torch.cuda.synchronize()
ts = []
for n in (1, 1 + iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    post, var, ls = be.vgp_fit(X, ym, yv, n, want_scale_tri=False)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print({"members": C * M, "T": T, "ms_per_iteration": (ts[1] - ts[0]) / iters, "ms_fixed": ts[0],
       "member_iterations_per_sec": C * M * iters / (ts[1] - ts[0]) * 1e3, "info": int(post.info_fit.max())})
