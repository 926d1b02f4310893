"""hsmc_b200 -- B200-native hard-sphere Monte Carlo hot path (drop-in for fedluc/HSMC's).

The product is the C-ABI shared library ``csrc/libhsmc_gpu.so`` (include/hsmc_gpu.h) and
the C host driver under ``host/``.  This Python package is plumbing around the C ABI for
tests, the benchmark and multi-process launches (torch.distributed); it contains no
compute and no CPU fallback -- if the CUDA library is missing or no GPU is visible every
entry point raises.
"""
from .gpu import HsmcGpu, HsmcError, load_library, library_path, ABI_SYMBOLS  # noqa: F401

__all__ = ["HsmcGpu", "HsmcError", "load_library", "library_path", "ABI_SYMBOLS"]
