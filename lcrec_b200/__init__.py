"""lcrec_b200 - B200-native (sm_100a) drop-in for LC-Rec's learned item-indexing hot path.

Importing the package never touches CUDA; the C-ABI library ``liblcrec_b200.so`` is loaded on the
first operator call and its absence is a hard error (there is no CPU fallback).
"""
__version__ = "0.1.0"
