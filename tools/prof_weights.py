import sys, json, torch
sys.path.insert(0, '.')
from bayesian_ensembling_b200.backend import Backend
be = Backend.get()
M, Ro, T = 24, 10, 1980
C = 4000000 // T
g = torch.Generator(device=be.device).manual_seed(1)
rnd = lambda *s: torch.rand(*s, dtype=torch.float64, device=be.device, generator=g)
a2 = 0.5 + rnd(C * M)
stats = torch.stack([a2, a2 * (0.9 + 0.2 * rnd(C * M)), a2 * (1.0 + 0.2 * rnd(C * M)), -0.5 * T * 1.8378770664093453 + rnd(C * M)], dim=1).contiguous()
obs = 0.8 + 0.4 * rnd(C, Ro, T)
if len(sys.argv) > 1 and sys.argv[1] == "time":
    best = 1e9
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); w = be.loglik_weights_mvn(stats, obs, M); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(best, bool(torch.isfinite(w).all()), float((w.sum(dim=1) - 1).abs().max()))
else:
    for _ in range(3):
        w = be.loglik_weights_mvn(stats, obs, M)
    torch.cuda.synchronize()
    print("ok")
