"""dot_ring_b200: B200-native (sm_100a) ring-proof engine behind dot-ring's Python API.

Host code is Python and calls the CUDA library ``libdotring_b200.so`` through ctypes
(``dot_ring_b200._native``).  The library is loaded on first use and its absence is a hard
``ImportError``: there is no CPU fallback.
"""

__all__ = ["__version__"]
__version__ = "0.1.0"
