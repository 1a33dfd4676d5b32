"""ensembles/utils.py entry points kept by the mirror (checkpoint loading only; ``PerfectModelTest`` is
orchestration and out of scope)."""
from .checkpoint import load_model_collection, load_reference_pickle  # noqa: F401
