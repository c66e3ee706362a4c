"""markovflow_b200: B200-native (sm_100a) implementation of Markovflow's structured linear-algebra
hot path behind the reference's operator API.  GPU only -- there is no CPU fallback."""
from ._lib import CholeskyError, MarkovflowB200Error
from .block_tri_diag import (
    BlockTriDiagonal,
    LowerTriangularBlockTriDiagonal,
    SymmetricBlockTriDiagonal,
)
from .config import set_check_numerics

__all__ = [
    "BlockTriDiagonal",
    "LowerTriangularBlockTriDiagonal",
    "SymmetricBlockTriDiagonal",
    "CholeskyError",
    "MarkovflowB200Error",
    "set_check_numerics",
]
