"""The oracle is test infrastructure: nothing under bayesian_ensembling_b200/ may import it,
and the product has no NumPy/SciPy linear-algebra fallback."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bayesian_ensembling_b200")


def _sources():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(dirpath, f)


def test_product_never_imports_oracle():
    for path in _sources():
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
        assert "reference_path" not in src, path


def test_product_has_no_cpu_linear_algebra():
    banned = ("np.linalg.", "numpy.linalg", "scipy.linalg", "torch.linalg", "torch.cholesky", "solve_triangular",
              "torch.matmul", "torch.bmm", "cusolver", "cublas")
    for path in _sources():
        src = open(path).read().replace("jnp.linalg", "")  # docstrings cite the reference's JAX calls
        for b in banned:
            assert b not in src, f"{b} in {path}"
