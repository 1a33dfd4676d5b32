"""Times be_dgemm_nt_i8tc (the library entry: device slicing + int8 tensor-core GEMM) against torch's fp64 matmul
(cuBLAS DGEMM) at a few shapes (developer tool).  usage: prof_i8tc.py [M N K]..."""
import json
import sys

import torch

sys.path.insert(0, '.')
from bayesian_ensembling_b200.backend import Backend

be = Backend.get()
shapes = [(3072, 3072, 3072), (2048, 4096, 3072), (4096, 4096, 1024)]
if len(sys.argv) > 3:
    shapes = [tuple(int(v) for v in sys.argv[1:4])]
g = torch.Generator(device=be.device).manual_seed(0)
for M, N, K in shapes:
    A = torch.randn(M, K, dtype=torch.float64, device=be.device, generator=g)
    B = torch.randn(N, K, dtype=torch.float64, device=be.device, generator=g)

    def timed(fn, reps=4):
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, out

    be.profile(True)
    be.profile_reset()
    ms_tc, C = timed(lambda: be.dgemm_nt_i8tc(A, B))
    prof = be.profile_read()
    be.profile(False)
    ms_cublas, Cref = timed(lambda: A @ B.T)
    err = float((C - Cref).abs().max() / (A.abs() @ B.abs().T).max())
    flops = 2.0 * M * N * K
    print(json.dumps({"M": M, "N": N, "K": K, "i8tc_ms_incl_slicing": ms_tc, "i8tc_tflops_incl_slicing": flops / ms_tc / 1e9,
                      "gemm_kernel_ms": prof["k_gemm_nt"]["ms"] / prof["k_gemm_nt"]["launches"],
                      "gemm_kernel_tflops": flops / (prof["k_gemm_nt"]["ms"] / prof["k_gemm_nt"]["launches"]) / 1e9,
                      "cublas_dgemm_ms": ms_cublas, "cublas_dgemm_tflops": flops / ms_cublas / 1e9,
                      "max_abs_diff_vs_cublas_over_max_sum_abs": err}))
