"""B200-native FastSVC generator forward (hand-written sm_100a CUDA behind a C ABI).

Light on import: the CUDA library is loaded lazily by ``svcc23_fastsvc_b200.abi``.
"""

__all__ = ["synthetic"]
