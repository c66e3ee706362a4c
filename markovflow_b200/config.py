"""Process-wide switches."""
_CHECK_NUMERICS = True


def set_check_numerics(flag: bool) -> None:
    """When True (default) factorisations synchronise on their per-chain ``info`` words and raise
    :class:`markovflow_b200.CholeskyError` on a non-positive pivot, like the reference's eager
    "Banded Cholesky decomposition failure".  Set False to stay fully asynchronous (NaNs propagate)."""
    global _CHECK_NUMERICS
    _CHECK_NUMERICS = bool(flag)


def check_numerics() -> bool:
    return _CHECK_NUMERICS
