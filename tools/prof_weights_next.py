"""CRPS / KSD / similarity weight kernels at bench.py's hbm_stages size, for an ncu capture (developer tool)."""
import sys

import torch

sys.path.insert(0, ".")
from bayesian_ensembling_b200.backend import Backend  # noqa: E402

be = Backend.get()
M, Ro, T = 24, 10, 1980
C = 4_000_000 // T
g = torch.Generator(device=be.device).manual_seed(1)
rnd = lambda *s: torch.rand(*s, dtype=torch.float64, device=be.device, generator=g)  # noqa: E731
obs = 0.8 + 0.4 * rnd(C, Ro, T)
means, variances = rnd(C, M, T), 0.01 + 0.05 * rnd(C, M, T)
sd = variances.sqrt()
for _ in range(2):
    be.crps_weights(means, sd, obs)
    be.ksd_weights(means, sd, obs)
    be.similarity_weights_pointwise(means, variances)
torch.cuda.synchronize()
print("ok")
