import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from bayesian_ensembling_b200.backend import Backend
from oracle import reference_path as rp
from test_gpu_parity import _posterior_covs
be = Backend.get()
T = 3
mus, covs = _posterior_covs(4, 5, T, seed=100 + T)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=be.device)
for idx1, idx2 in (([0, 1, 2, 3], [1, 2, 3, 0]), ([0, 1, 2, 3, 1], [1, 2, 3, 0, 1]), ([1], [1]), ([0, 1, 2, 3, 0], [1, 2, 3, 0, 1])):
    S1, S2 = covs[idx1], covs[idx2]
    R1, _, it1, _ = be.sqrtm_psd(t(S1))
    R1 = R1.cpu().numpy()
    H = np.stack([R1[i] @ S2[i] @ R1[i] for i in range(len(idx1))])
    H = 0.5 * (H + H.transpose(0, 2, 1))
    for mi in (40, 8, 12, 20):
        Q, _, it2, info = be.sqrtm_psd(t(H), max_iters=mi)
        Q = Q.cpu().numpy()
        print(idx1, "max", mi, "iters", it1, it2, "sqrt(H) err", [float(np.abs(Q[i] - rp.sqrtm_svd(H[i])).max() / np.abs(Q[i]).max()) for i in range(len(idx1))])
    w2, _ = be.w2_distance(t(mus[idx1]), t(S1), t(mus[idx2]), t(S2))
    want = [rp.gaussian_w2_distance(mus[a], covs[a], mus[b], covs[b]) for a, b in zip(idx1, idx2)]
    print("  w2 err", np.abs(w2.cpu().numpy() - want))
