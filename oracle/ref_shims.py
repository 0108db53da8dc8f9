"""Import the UNMODIFIED reference (`/root/reference`) in a container that lacks some of
its third-party imports.

TEST INFRASTRUCTURE ONLY.  This module is used by `oracle/make_golden.py` (run in the
build container, where `/root/reference` is mounted) to generate the committed golden
fixtures under `tests/golden/`.  Nothing in the product package, `bench.py`, `smoke()` or
the `-m gpu` tests imports it: `/root/reference` does not exist on the GPU box.

The reference fails to import only because of module-level imports of packages that are
not installed here and that the EDM / sampler path never calls
(`fastgen/utils/__init__.py:12` omegaconf, `fastgen/utils/basic_utils.py:13,19` ftfy /
imageio, `fastgen/networks/noise_schedule.py:13` diffusers, `fastgen/utils/io_utils.py:5`
boto3).  We register inert stand-ins for exactly those names; every arithmetic line that
runs is the reference's own.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("FASTGEN_REF", "/root/reference")


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _AttrDict(dict):
    """Stand-in for omegaconf.DictConfig: attribute access over a dict."""

    def __init__(self, content=None, flags=None, **kw):
        super().__init__()
        if content:
            self.update(content)
        self.update(kw)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class _ListConfig(list):
    pass


class _OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True, **kw):
        return cfg

    @staticmethod
    def create(obj=None, **kw):
        return _AttrDict(obj or {})

    @staticmethod
    def is_config(obj):
        return isinstance(obj, (_AttrDict, _ListConfig))


def install_shims() -> None:
    """Register the stand-in modules (idempotent; real packages win if installed)."""
    def missing(name):
        if name in sys.modules:
            return False
        try:
            importlib.import_module(name)
            return False
        except Exception:
            return True

    if missing("omegaconf"):
        _module("omegaconf", DictConfig=_AttrDict, ListConfig=_ListConfig, OmegaConf=_OmegaConf)
    if missing("ftfy"):
        _module("ftfy", fix_text=lambda s: s)
    if missing("imageio"):
        v3 = _module("imageio.v3")
        _module("imageio", v3=v3)
    if missing("boto3"):
        _module("boto3")
    if missing("diffusers"):
        _module("diffusers", DDIMScheduler=type("DDIMScheduler", (), {}),
                CogVideoXDPMScheduler=type("CogVideoXDPMScheduler", (), {}))


def load_reference():
    """Return the imported reference package `fastgen` (raises if the tree is absent)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "fastgen")):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("fastgen")
