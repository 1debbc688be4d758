from . import linen  # noqa: F401
