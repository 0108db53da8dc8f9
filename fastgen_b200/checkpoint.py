"""Checkpoint I/O compatible with the reference -- SURVEY.md 8(f) item 1.

* ``save_checkpoint`` / ``load_checkpoint`` read and write the ``.pth`` layout of ``fastgen.utils.checkpointer.Checkpointer``
  (reference utils/checkpointer.py:58-111 save, :113-200 load): one ``torch.save`` dict ``{"model": {name: state_dict},
  "optimizer", "scheduler", "grad_scaler", "callbacks", "iteration"}``; inference loads the ``"net"`` entry and the EMA
  entries with ``strict=False`` exactly like ``scripts/inference/inference_utils.py:117-136``.  Our networks keep the
  reference's ``state_dict`` key names, so a checkpoint written by either side loads on the other.
* ``load_safetensors`` loads Hugging Face ``*.safetensors`` weights (a single file, or a directory with
  ``diffusion_pytorch_model.safetensors`` / a sharded ``*.safetensors.index.json``) into a sub-module, e.g. the diffusers
  ``transformer/`` folder of Wan 2.1 under the ``transformer.`` prefix or the SDXL ``unet/`` folder under ``unet.``
  (what ``from_pretrained`` does for the reference, Wan/network.py:659-681, SDXL/network.py:210-230).

Host-side only: tensors are loaded to CPU (or ``device``) and copied by ``load_state_dict``; the CUDA layouts are rebuilt
lazily on the next forward (every network invalidates its packed weights in ``load_state_dict``).
"""
from __future__ import annotations

import json
import os
import warnings
from typing import Any, Dict, Mapping, Optional

import torch

__all__ = ["save_checkpoint", "load_checkpoint", "load_safetensors"]


def save_checkpoint(model_dict: Mapping[str, torch.nn.Module], path: str, iteration: int = 0) -> str:
    """Write ``model_dict`` (e.g. ``{"net": net, "ema": ema_net}``) in the reference's ``.pth`` layout."""
    assert path.endswith(".pth"), f"{path} does not end with .pth"
    save_dict = {
        "model": {k: v.state_dict() for k, v in model_dict.items()},
        "optimizer": None, "scheduler": None, "grad_scaler": None, "callbacks": None,
        "iteration": iteration,
    }
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(save_dict, path)
    return path


def load_checkpoint(model_dict: Mapping[str, Optional[torch.nn.Module]], path: str, device: Any = "cpu") -> int:
    """Load the ``"model"`` entries of a reference checkpoint into the modules of ``model_dict`` (``strict=False``, like
    the reference, so external checkpoints without e.g. ``logvar_linear`` load); returns the stored iteration, 0 when
    the file does not exist or holds none."""
    assert path.endswith(".pth"), f"{path} does not end with .pth"
    if not os.path.exists(path):
        warnings.warn(f"Checkpoint file not found at {path}")
        return 0
    state = torch.load(path, map_location=device, weights_only=False)
    for k, v in model_dict.items():
        if k in state["model"] and v is not None:
            info = v.load_state_dict(state["model"][k], strict=False)
            if info.unexpected_keys:
                warnings.warn(f"Model {k}: unexpected keys {list(info.unexpected_keys)[:5]} ...")
        else:
            warnings.warn(f"Model {k} not found in checkpoint.")
    return int(state.get("iteration", 0) or 0)


def _safetensors_files(path: str):
    if os.path.isfile(path):
        return [path]
    index = [f for f in sorted(os.listdir(path)) if f.endswith(".safetensors.index.json")]
    if index:
        with open(os.path.join(path, index[0])) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
        return [os.path.join(path, s) for s in shards]
    files = [os.path.join(path, f) for f in sorted(os.listdir(path)) if f.endswith(".safetensors")]
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    return files


def load_safetensors(module: torch.nn.Module, path: str, prefix: str = "", strict: bool = True,
                     device: str = "cpu") -> Dict[str, Any]:
    """Load Hugging Face safetensors weights; every key is prefixed with ``prefix`` (``"transformer."`` for a diffusers Wan
    transformer folder, ``"unet."`` for an SDXL UNet folder).  With ``strict`` every parameter below ``prefix`` must be
    present in the files, except the modules FastGen itself adds (``logvar_linear``, ``r_embedder``); parameters outside
    the prefix are left alone."""
    from safetensors.torch import load_file

    sd: Dict[str, torch.Tensor] = {}
    for f in _safetensors_files(path):
        for k, v in load_file(f, device=device).items():
            sd[prefix + k] = v
    info = module.load_state_dict(sd, strict=False)
    # modules FastGen adds on top of the Hugging Face model are never in its files (logvar_linear: Wan/network.py:640,
    # SDXL/network.py:231; the r embedder: Wan/network.py:615-617)
    extras = ("logvar_linear.", "r_embedder.")
    missing = [k for k in info.missing_keys if k.startswith(prefix) and not any(e in k for e in extras)]
    if strict and (missing or info.unexpected_keys):
        raise RuntimeError(f"load_safetensors: missing {missing[:5]} (+{max(len(missing) - 5, 0)}), "
                           f"unexpected {list(info.unexpected_keys)[:5]}")
    return {"missing": missing, "unexpected": list(info.unexpected_keys), "tensors": len(sd)}
