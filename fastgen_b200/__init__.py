"""fastgen_b200 -- B200-native (sm_100a) few-step student sampler for FastGen networks.

The compute path is hand-written CUDA behind a C ABI (``include/fastgen_b200.h``); this package holds the
ctypes binding (``_lib``), tensor-level wrappers (``ops``) and the host-side mirrors of the reference's
``networks/`` forward contract and ``methods/`` sampler API.

The public classes are importable from the package root (resolved lazily, so ``import fastgen_b200`` stays cheap and does
not need the built library):

    networks   EDMPrecond, Wan, CausalWan, StableDiffusionXL, DiT            (fastgen/networks/*)
    samplers   FastGenModel, MeanFlowModel, CausVidModel, SelfForcingModel     (fastgen/methods/*)
    schedules  get_noise_schedule                                             (fastgen/networks/noise_schedule.py)
    I/O        load_checkpoint, save_checkpoint, load_safetensors              (fastgen/utils/checkpointer.py)
"""
__version__ = "0.1.0"

_EXPORTS = {
    "EDMPrecond": "edm", "Wan": "wan", "CausalWan": "wan", "StableDiffusionXL": "sdxl", "DiT": "dit",
    "FastGenModel": "sampler", "MeanFlowModel": "sampler", "CausVidModel": "sampler", "SelfForcingModel": "sampler",
    "get_noise_schedule": "noise_schedule",
    "load_checkpoint": "checkpoint", "save_checkpoint": "checkpoint", "load_safetensors": "checkpoint",
    "FastGenB200Error": "_lib",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module 'fastgen_b200' has no attribute '{name}'")
    import importlib

    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)
