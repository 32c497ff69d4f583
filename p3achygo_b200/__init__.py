"""p3achygo_b200 — B200-native (sm_100a) batched leaf evaluation behind p3achygo's ``nn::Engine``.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of include/p3_b200.h),
``engine.py`` (host mirror of the reference engine interface over that ABI), ``weights.py`` (flat
weight files + net shapes) and ``host/`` (the C++ adapter / benchmark harness).
Importing ``p3achygo_b200.engine`` requires the built shared library; there is no fallback path.
"""
__version__ = "0.1"
