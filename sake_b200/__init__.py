"""sake_b200 — B200-native (sm_100a) DenseSAKELayer / DenseSAKEModel behind the reference's
flax-style module API (sake/layers.py, sake/models.py).  The compute path is the CUDA library
`libsake_b200.so` (C ABI in include/sake_b200.h); importing this package without the built
library raises — there is no CPU fallback."""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from . import functional, utils, layers, models, flows  # noqa: F401
from .layers import DenseSAKELayer  # noqa: F401
from .models import DenseSAKEModel  # noqa: F401

__all__ = ["DenseSAKELayer", "DenseSAKEModel", "functional", "utils", "layers", "models", "flows"]
