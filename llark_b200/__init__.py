"""llark_b200: B200-native (sm_100a) implementation of the LLark audio-conditioned training hot path.

The package mirrors the reference's Python surface for that path (m2t/models/llamav2.py, m2t/data_modules.py,
m2t/models/trainer.py, jukebox/main.py) and executes it through hand-written CUDA behind the C-ABI in
include/llark_b200.h. Importing the package does not need a GPU; calling any op does (no CPU fallback).
"""
__version__ = "0.1.0"
