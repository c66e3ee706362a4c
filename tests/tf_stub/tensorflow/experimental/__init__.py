from . import dlpack  # noqa: F401
