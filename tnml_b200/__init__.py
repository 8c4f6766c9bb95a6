"""tnml_b200 -- B200-native `fixedL` bond-update path of emstoudenmire/TNML.

The product is the CUDA library `libtnml_b200.so` behind the C-ABI of
`include/tnml_b200.h`; this package is its thin Python host side
(`capi` = ctypes binding, `fixedl` = mirror of the reference's
TrainStates / cgrad / quadcost / mldmrg interface, `data` = MNIST idx reader,
feature map and input-file parser).  Nothing here computes on the CPU.
"""
from . import capi  # noqa: F401

__all__ = ["capi", "fixedl", "data"]
