"""krig-b200: B200-native (sm_100a) implementation of KADAL's Kriging hot path.

Drop-in for the reference's surrogate API on that path only:
``surrogate_models`` (Kriging / KPLS, likelihood, calckernel, prediction),
``optim_tools`` (SOBO EI sweep) and ``reliability_analysis`` (AK-MCS).
All arithmetic runs in hand-written CUDA kernels behind ``libkrigb200.so``
(``include/krigb200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
