"""deep_rawburst_sr_b200 -- B200-native (sm_100a) implementation of the DBSR burst forward pass, drop-in for the
DBSRNet / PWCNet / WeightedSum / correlation.FunctionCorrelation interfaces of Tony-Tseng/deep-rawburst-sr.

The compute path is libdbsr_b200.so (hand-written CUDA behind the C ABI of include/dbsr_b200.h); importing the
package does not need a GPU, running any op does (no CPU fallback).
"""
__version__ = '0.1.0'

from ._lib import LIB_PATH, DbsrB200Error, load_library  # noqa: F401
