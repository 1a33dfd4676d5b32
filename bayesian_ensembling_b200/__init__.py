"""B200-native fit -> weight -> barycentre hot path of mattramos/bayesian_ensembling.

Mirrors the reference's public names (ensembles/__init__.py:1-6) for the path in scope."""
from .data import Distribution, ModelCollection, ProcessModel  # noqa: F401
from .dtw import dtw_barycenter_averaging_subgradient, performDBA  # noqa: F401
from .ensemble_scheme import Barycentre  # noqa: F401
from .labelled import DataArray  # noqa: F401
from .models import GPDTW1D, GPDTW3D, MeanFieldApproximation  # noqa: F401
from .wasserstein import (gaussian_barycentre, gaussian_barycentre_fullcov, gaussian_w2_distance_distrax,  # noqa: F401
                          sqrtm, wasserstien_distance)
from .weights import (CRPSWeight, InverseSquareWeight, KSDWeight, LogLikelihoodWeight, ModelSimilarityWeight,  # noqa: F401
                      UniformWeight)

__version__ = "0.1.0"
