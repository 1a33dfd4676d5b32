"""A few SVGP optimisation steps at the reference's sizes (M = 400, minibatch 500), for an ncu launch list (developer tool)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bayesian_ensembling_b200.backend import Backend  # noqa: E402
from bayesian_ensembling_b200.models import GPDTW3D  # noqa: E402

be = Backend.get()
rng = np.random.default_rng(1)
N, R, M, batch, steps = 20000, 10, 400, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 3
lat, lon, t = rng.uniform(-85, 85, N), rng.uniform(0, 360, N), rng.uniform(-1, 1, N)
X = np.column_stack([np.cos(np.radians(lat)) * np.cos(np.radians(lon)), np.cos(np.radians(lat)) * np.sin(np.radians(lon)),
                     np.sin(np.radians(lat)), t, 0.4 * t[:, None] + 0.1 * rng.standard_normal((N, R))])
Y = np.column_stack([0.4 * t + 0.2 * np.sin(np.radians(lat)) + 0.05 * rng.standard_normal(N), rng.uniform(0.005, 0.03, N)])
Z0 = np.linspace(X.min(axis=0), X.max(axis=0), M)
idx = GPDTW3D.minibatch_order(N, batch, 2 * steps, 1)
Xd, Yd, Zd = (torch.as_tensor(a, device=be.device) for a in (X, Y, Z0))
out = be.svgp_fit(Xd, Yd, Zd, idx, steps)
torch.cuda.synchronize()
print("ok", int(out["info"].item()))
