"""Measures the FP64 denominators on the GPU box: cuBLAS DGEMM (burst + sustained) and,
for context only, cuSOLVER batched potrf at the cfg2 shape.  Library calls here size the
roofline; they are never on the product path."""
import json, sys, time
import torch

def ev_time(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {"gpu": torch.cuda.get_device_name(0)}
torch.backends.cuda.matmul.allow_tf32 = False
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    ms = ev_time(lambda: torch.matmul(a, b, out=c))
    out[f"dgemm_{n}_tflops_burst"] = 2 * n**3 / ms / 1e9
n = 8192
t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5): torch.matmul(a, b, out=c)
    k += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_tflops_sustained"] = 2 * n**3 * k / e0.elapsed_time(e1) / 1e9
# NT form (syrk-like) 
ms = ev_time(lambda: torch.matmul(a, b.T, out=c)); out["dgemm_8192_nt_tflops"] = 2 * n**3 / ms / 1e9
del a, b, c
# context: cuSOLVER batched Cholesky at cfg2 shape
T, B = 3012, 24
x = torch.randn(B, T, 64, dtype=torch.float64, device="cuda")
A = x @ x.transpose(1, 2) + 64 * torch.eye(T, dtype=torch.float64, device="cuda")
ms = ev_time(lambda: torch.linalg.cholesky(A), reps=3)
out["cusolver_potrf_24x3012_ms"] = ms
out["cusolver_potrf_24x3012_tflops"] = B * T**3 / 3 / ms / 1e9
ms = ev_time(lambda: torch.linalg.cholesky(A[0]), reps=3)
out["cusolver_potrf_1x3012_ms"] = ms
# hbm copy
src = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
ms = ev_time(lambda: dst.copy_(src)); out["copy_gbs"] = 2 * src.numel() * 8 / ms / 1e6
print(json.dumps(out))
