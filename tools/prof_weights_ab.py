"""k_loglik_weights_mvn_tab against the library-exp form: accuracy of exp_tab16 over its whole range, the
special cases of the normaliser, and the time of both forms at 4 M points.  Run twice: plainly and with
BE_WEIGHTS_LIBEXP=1 (the A/B switch in be_loglik_weights_mvn)."""
import sys, json, os, torch
sys.path.insert(0, '.')
from bayesian_ensembling_b200.backend import Backend
be = Backend.get()
dev = be.device
out = {"tab": not os.environ.get("BE_WEIGHTS_LIBEXP")}
# --- accuracy: stats chosen so that cst * mean sweeps [-760, 760] (a = ab = 0, bb = 0, logdet = -x - T/2 log 2pi)
M, Ro, T = 24, 3, 4096
C = 8
g = torch.Generator(device=dev).manual_seed(3)
x = (torch.rand(C * M, dtype=torch.float64, device=dev, generator=g) - 0.5) * 1520.0
x[0], x[1], x[2], x[3] = 0.0, 709.7, -745.0, 699.9999
stats = torch.zeros(C * M, 4, dtype=torch.float64, device=dev)
stats[:, 0] = 1e-3  # a small quadratic term so that the points of a row differ
stats[:, 3] = -x - 0.5 * T * 1.8378770664093453
obs = torch.rand(C, Ro, T, dtype=torch.float64, device=dev, generator=g) * 40.0
w, le, lm = be.loglik_weights_mvn(stats, obs, M, want_lls=True)
w2 = be.loglik_weights_mvn(stats, obs, M)
ref = torch.exp(lm)
fin = torch.isfinite(ref) & (ref > 1e-300)
rel = ((le - ref).abs() / ref)[fin]
out["exp_max_rel_err_vs_torch"] = float(rel.max())
slow = ~(lm.abs() < 700)  # the arguments that take the library's exp in both kernels
out["exp_special_equal"] = bool(torch.equal(le[slow].nan_to_num(nan=-1.0), ref[slow].nan_to_num(nan=-1.0)))
out["n_fast"] = int((lm.abs() < 700).sum()); out["n_slow"] = int((lm.abs() >= 700).sum())
tot = le.sum(dim=1, keepdim=True)
wref = le / tot
okw = torch.isfinite(wref)
out["w_max_abs_err_vs_division"] = float((w - wref)[okw].abs().max())
out["w_nan_pattern_equal"] = bool(torch.equal(torch.isnan(w), torch.isnan(wref)))
out["w_lls_vs_plain_equal"] = bool(torch.equal(w.nan_to_num(nan=-1.0), w2.nan_to_num(nan=-1.0)))
# --- normaliser special cases: all-underflow (0/0), subnormal total, overflow (inf/inf), NaN
for name, xv in (("all_zero", -800.0), ("subnormal_total", -740.0), ("overflow", 720.0), ("nan", float("nan"))):
    s2 = torch.zeros(M, 4, dtype=torch.float64, device=dev)
    s2[:, 3] = -xv - 0.5 * 64 * 1.8378770664093453
    o2 = torch.zeros(1, 2, 64, dtype=torch.float64, device=dev)
    w_, le_, lm_ = be.loglik_weights_mvn(s2, o2, M, want_lls=True)
    r_ = torch.exp(lm_); r_ = r_ / r_.sum(dim=1, keepdim=True)
    out["special_" + name] = bool(torch.allclose(w_, r_, rtol=1e-14, atol=0, equal_nan=True))
# --- time at 4 M points (tools/prof_weights.py's operands)
M, Ro, T = 24, 10, 1980
C = 4000000 // T
rnd = lambda *s: torch.rand(*s, dtype=torch.float64, device=dev, generator=g)
a2 = 0.5 + rnd(C * M)
stats = torch.stack([a2, a2 * (0.9 + 0.2 * rnd(C * M)), a2 * (1.0 + 0.2 * rnd(C * M)), -0.5 * T * 1.8378770664093453 + rnd(C * M)], dim=1).contiguous()
obs = 0.8 + 0.4 * rnd(C, Ro, T)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w = be.loglik_weights_mvn(stats, obs, M); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[2:])
nbytes = (C * Ro * T + C * M * 4 + C * M * T) * 8
out["ms_median"] = ts[len(ts) // 2]; out["ms_min"] = ts[0]
out["GBps_median"] = nbytes / out["ms_median"] / 1e6
out["finite"] = bool(torch.isfinite(w).all()); out["sum_err"] = float((w.sum(dim=1) - 1).abs().max())
print(json.dumps(out))
