"""Times the device DBA (be_dtw_barycenter_averaging_subgradient) on synthetic cells of a BASELINE
shape with the per-kernel profiler (developer tool; also the target of ncu captures).
usage: prof_dba.py [cfg=cfg2] [cells=6] [max_iter=50] [reps=2]"""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
from bayesian_ensembling_b200 import synthetic
from bayesian_ensembling_b200.backend import Backend

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 6
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 50
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
be = Backend.get()
cfg = synthetic.CONFIGS[name]
r, _ = synthetic.make_cells(cfg, n_cells=cells)
reals = torch.as_tensor(r, device=be.device)
C, M, R, T = reals.shape
X = reals.reshape(C * M, R, T).contiguous()
out = {}
for rep in range(reps + 1):
    be.profile(True)
    be.profile_reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bary, n_iter, cost = be.dtw_barycenter_averaging_subgradient(X, max_iter=max_iter, tol=1e-3, want_info=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = be.profile_read()
    be.profile(False)
    out = dict(cfg=name, cells=C, members=M, R=R, T=T, max_iter=max_iter, ms=ms, cells_per_s=C / ms * 1e3,
               n_iter_min=int(n_iter.min()), n_iter_max=int(n_iter.max()), n_iter_mean=float(n_iter.float().mean()),
               kernels={k: v for k, v in prof.items() if k.startswith("k_dtw") or k.startswith("k_dba")})
for k, v in out["kernels"].items():
    v["tflops"] = v["flops"] / v["ms"] / 1e9 if v["ms"] else 0.0
    v["gcells_per_s"] = v["flops"] / 5.0 / v["ms"] / 1e6 if v["ms"] else 0.0
print(json.dumps(out))
