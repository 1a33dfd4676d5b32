"""ensembles/utils.py entry points kept by the mirror: checkpoint loading and ``PerfectModelTest``."""
from .checkpoint import load_model_collection, load_reference_pickle  # noqa: F401
from .perfect_model import PerfectModelTest  # noqa: F401
