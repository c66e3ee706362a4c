"""markovflow_b200: B200-native (sm_100a) implementation of Markovflow's structured linear-algebra
hot path behind the reference's operator API.  GPU only -- there is no CPU fallback."""
from ._lib import CholeskyError, MarkovflowB200Error
from .block_tri_diag import (
    BlockTriDiagonal,
    LowerTriangularBlockTriDiagonal,
    SymmetricBlockTriDiagonal,
)
from .config import set_check_numerics
from .emission_model import EmissionModel
from .gauss_markov import GaussMarkovDistribution, check_compatible
from .kalman_filter import (
    BaseKalmanFilter,
    GaussianSites,
    KalmanFilter,
    KalmanFilterWithSites,
    KalmanFilterWithSparseSites,
    UnivariateGaussianSitesNat,
    kalman_log_likelihood,
)
from .state_space_model import (
    StateSpaceModel,
    cholesky_or_zero,
    state_space_model_from_covariances,
)

__all__ = [
    "BlockTriDiagonal",
    "LowerTriangularBlockTriDiagonal",
    "SymmetricBlockTriDiagonal",
    "CholeskyError",
    "MarkovflowB200Error",
    "set_check_numerics",
    "EmissionModel",
    "GaussMarkovDistribution",
    "check_compatible",
    "BaseKalmanFilter",
    "GaussianSites",
    "KalmanFilter",
    "KalmanFilterWithSites",
    "KalmanFilterWithSparseSites",
    "UnivariateGaussianSitesNat",
    "kalman_log_likelihood",
    "StateSpaceModel",
    "cholesky_or_zero",
    "state_space_model_from_covariances",
]
