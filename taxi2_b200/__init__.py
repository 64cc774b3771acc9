"""taxi2_b200 -- B200-native implementation of TaxI2's pairwise-distance hot path.

Python host code over a C-ABI CUDA library (taxi2_b200/lib/libtaxi2_b200.so, sm_100a).
There is no CPU fallback: compute entry points raise when the library or a GPU is missing.
"""
__version__ = "0.1.0"
