"""tnqvm_b200 -- B200-native matrix-product-state gate engine behind TNQVM's `exatn-mps` visitor surface.

The product path is CUDA only (libmps_b200.so, sm_100a).  Importing this package never imports any CPU
checker; constructing an engine without the built library or without a GPU raises.
"""
from .abi import lib_path, load_library, MpsError  # noqa: F401
from .mps import B200MPS, CompiledCircuit  # noqa: F401
from . import circuits, gates  # noqa: F401
