"""meshclust_b200 -- B200-native (sm_100a) implementation of the MeShClust data-parallel hot path.

The product is the CUDA library ``libmeshclust_b200.so`` (C-ABI in ``include/meshclust_b200.h``)
and the host CLI ``bin/meshclust``.  This Python package is a thin ctypes mirror of the C-ABI used
by the tests and the benchmark; it never computes anything itself and there is no CPU fallback:
importing :mod:`meshclust_b200.api` fails loudly when the library has not been built, and every
compute call fails with ``MC_ERR_CUDA`` when no GPU is present.
"""

__all__ = ["api", "synth", "build"]
__version__ = "0.1"
