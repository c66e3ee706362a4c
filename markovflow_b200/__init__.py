"""markovflow_b200: B200-native (sm_100a) implementation of Markovflow's structured linear-algebra
hot path behind the reference's operator API.  GPU only -- there is no CPU fallback."""
from ._lib import CholeskyError, MarkovflowB200Error
from .block_tri_diag import (
    BlockTriDiagonal,
    LowerTriangularBlockTriDiagonal,
    SymmetricBlockTriDiagonal,
)
from .conditionals import (
    base_conditional_predict,
    conditional_predict,
    conditional_statistics,
    conditional_predict_from_transitions,
    conditional_statistics_from_transitions,
    insertion_indices,
    pairwise_marginals,
)
from .config import set_check_numerics
from .emission_model import EmissionModel
from .gauss_markov import GaussMarkovDistribution, check_compatible
from .graphs import Graphed
from .kalman_filter import (
    BaseKalmanFilter,
    GaussianSites,
    KalmanFilter,
    KalmanFilterWithSites,
    KalmanFilterWithSparseSites,
    UnivariateGaussianSitesNat,
    kalman_log_likelihood,
)
from .ssm_gaussian_transformations import (
    expectations_to_ssm_params,
    naturals_to_ssm_params,
    naturals_to_ssm_params_no_smoothing,
    ssm_to_expectations,
    ssm_to_naturals,
    ssm_to_naturals_no_smoothing,
)
from .ssm_natgrad import SSMNaturalGradient
from .kernels import HarmonicOscillator, Matern12, Matern32, Matern52, SDEKernel, Sum, matern_kalman_log_likelihood
from .posterior import (
    AnalyticPosteriorProcess,
    ConditionalProcess,
    ImportanceWeightedPosteriorProcess,
    PosteriorProcess,
)
from .state_space_model import (
    StateSpaceModel,
    cholesky_or_zero,
    state_space_model_from_covariances,
)

__all__ = [
    "Graphed",
    "AnalyticPosteriorProcess",
    "ConditionalProcess",
    "ImportanceWeightedPosteriorProcess",
    "PosteriorProcess",
    "HarmonicOscillator",
    "SDEKernel",
    "Sum",
    "conditional_predict",
    "conditional_statistics",
    "SSMNaturalGradient",
    "base_conditional_predict",
    "conditional_predict_from_transitions",
    "conditional_statistics_from_transitions",
    "insertion_indices",
    "pairwise_marginals",
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
    "Matern12",
    "Matern32",
    "Matern52",
    "matern_kalman_log_likelihood",
    "StateSpaceModel",
    "expectations_to_ssm_params",
    "naturals_to_ssm_params",
    "naturals_to_ssm_params_no_smoothing",
    "ssm_to_expectations",
    "ssm_to_naturals",
    "ssm_to_naturals_no_smoothing",
    "cholesky_or_zero",
    "state_space_model_from_covariances",
]
