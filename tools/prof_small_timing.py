"""Per-phase cycle counts of one CTA of the small-T member kernels (developer tool; needs the -DBE_SMALL_TIMING
library: tools/gpu_small_timing.sh)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bayesian_ensembling_b200 import grid, synthetic  # noqa: E402
from bayesian_ensembling_b200.backend import Backend  # noqa: E402

be = Backend.get()
cfg = synthetic.CONFIGS["cfg4"]
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reals, obs = synthetic.make_cells(cfg, n_cells=cells)
r = torch.as_tensor(reals, device=be.device)
o = torch.as_tensor(obs, device=be.device)
for _ in range(2):
    grid.fit_weight_barycentre(r, o, 0.5, 6.0, cells_per_wave=cells)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
be.lib.be_debug_small_timing(None, 1)
reps = 3
for _ in range(reps):
    grid.fit_weight_barycentre(r, o, 0.5, 6.0, cells_per_wave=cells)
torch.cuda.synchronize()
be.lib.be_debug_small_timing(buf, 0)
t = np.array(list(buf), dtype=np.float64).reshape(2, 16, 2, 8) / reps
names = ["window: D inverse / G update_early", "S1 (clock read may precede the wait)", "diag tile + scale (+ S1 wait)", "S2", "late update (+ S2 wait)", "S3", "D trailing+barriers / G extra", "D pivot chain"]
for kern, kn in enumerate(["k_small_factor_inverse", "k_small_cov_factor"]):
    print(f"== {kn}: cycles per phase of one CTA (rows: block step k; 'D' = diagonal group thread 0, 'G' = product group thread 128)")
    print("step | " + " | ".join(f"{n[:18]:>18s}" for n in names[:8]))
    tot = np.zeros((2, 8))
    for k in range(8):
        for g, gn in enumerate("DG"):
            print(f"{k} {gn}  | " + " | ".join(f"{t[kern, k, g, p]:18.0f}" for p in range(8)))
            tot[g] += t[kern, k, g]
    for g, gn in enumerate("DG"):
        print(f"sum {gn}| " + " | ".join(f"{tot[g, p]:18.0f}" for p in range(8)), f"  total {tot[g].sum():.0f}")
    print("pre / post:", {i: float(t[kern, 15, 0, i]) for i in range(4)})
