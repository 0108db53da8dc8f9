"""fastgen_b200 -- B200-native (sm_100a) few-step student sampler for FastGen networks.

The compute path is hand-written CUDA behind a C ABI (``include/fastgen_b200.h``); this package holds the
ctypes binding (``_lib``), tensor-level wrappers (``ops``) and the host-side mirrors of the reference's
``networks/`` forward contract and ``methods/`` sampler API.
"""
__version__ = "0.1.0"
