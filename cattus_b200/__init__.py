"""cattus_b200: B200-native (sm_100a) batched neural-network evaluation for the Cattus MCTS engine.

The product is the C-ABI shared library `libcattus_b200.so` (include/cattus_b200.h); this package holds its CUDA
sources (csrc/), the in-tree build recipe, the weight-blob exporter and a host-side mirror of the reference's
`NNetwork` interface over ctypes.  Importing the package does not load the library; constructing a `CudaNetwork`
does, and fails loudly without it or without an sm_100 device.
"""
from .cache import ValueFuncCache  # noqa: F401
from .export import export_blob, export_model  # noqa: F401
from .network import CudaNetwork  # noqa: F401

__all__ = ["CudaNetwork", "ValueFuncCache", "export_blob", "export_model"]
