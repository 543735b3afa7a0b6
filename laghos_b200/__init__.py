"""laghos_b200 — B200-native partial-assembly hot path of Laghos (C ABI + thin Python binding).

The product is the CUDA library ``laghos_b200/lib/liblaghos_b200.so`` (include/laghos_b200.h).
This package only binds it with ctypes for the tests, the benchmark and the multi-GPU
launcher; PyTorch supplies device memory, streams and ``torch.distributed``.
There is no CPU fallback: importing :mod:`laghos_b200.api` without the built library raises.
"""
from ._lib import load_library, LIB_PATH  # noqa: F401
