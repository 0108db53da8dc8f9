"""Checkpoint compatibility: a `.pth` in the reference Checkpointer's layout, written from a real
reference network (tests/golden/ref_checkpoint_tiny.pth, oracle/make_golden.py), loads into our modules tensor for
tensor; our writer produces the same layout; Hugging Face safetensors folders load under a prefix."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cfg():
    return json.loads(str(np.load(os.path.join(GOLD, "ref_checkpoint_tiny_io.npz"))["cfg"]))


def test_reference_checkpoint_loads_tensor_for_tensor():
    from fastgen_b200.checkpoint import load_checkpoint
    from fastgen_b200.edm import EDMPrecond

    path = os.path.join(GOLD, "ref_checkpoint_tiny.pth")
    state = torch.load(path, map_location="cpu", weights_only=False)
    assert set(state) == {"model", "optimizer", "scheduler", "grad_scaler", "callbacks", "iteration"}
    net, ema = EDMPrecond(**_cfg()), EDMPrecond(**_cfg())
    before = {k: v.clone() for k, v in ema.state_dict().items()}
    with pytest.warns(UserWarning, match="fake_score not found"):
        it = load_checkpoint({"net": net, "ema": ema, "fake_score": None}, path)
    assert it == 7
    sd = net.state_dict()
    assert set(sd) == set(state["model"]["net"])                       # same key names as the reference network
    for k, v in state["model"]["net"].items():
        assert torch.equal(sd[k], v), k
    # the EMA entry is partial: strict=False loads what is there and leaves the rest alone
    for k, v in ema.state_dict().items():
        want = state["model"]["ema"].get(k, before[k])
        assert torch.equal(v, want), k
    assert load_checkpoint({"net": net}, os.path.join(GOLD, "does_not_exist.pth")) == 0


def test_save_checkpoint_round_trip(tmp_path):
    from fastgen_b200.checkpoint import load_checkpoint, save_checkpoint
    from fastgen_b200.edm import EDMPrecond

    torch.manual_seed(0)
    a, b = EDMPrecond(**_cfg()), EDMPrecond(**_cfg())
    with torch.no_grad():
        for prm in a.parameters():
            prm.normal_()
    path = save_checkpoint({"net": a}, str(tmp_path / "ckpt" / "0000012.pth"), iteration=12)
    assert load_checkpoint({"net": b}, path) == 12
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k
    state = torch.load(path, weights_only=False)
    assert list(state["model"]) == ["net"] and state["optimizer"] is None


def test_safetensors_folder_loads_under_prefix(tmp_path):
    """a diffusers-style transformer folder (sharded safetensors + index) into Wan's `transformer.` sub-module"""
    from safetensors.torch import save_file

    from fastgen_b200.checkpoint import load_safetensors
    from fastgen_b200.wan import Wan
    from oracle import wan_oracle as WO

    cfg = WO.tiny_config(layers=2)
    sd = WO.make_state_dict(cfg, 3)
    inner = {k[len("transformer."):]: v.contiguous() for k, v in sd.items()}
    keys = sorted(inner)
    folder = tmp_path / "transformer"
    folder.mkdir()
    shards = {"diffusion_pytorch_model-00001-of-00002.safetensors": keys[: len(keys) // 2],
              "diffusion_pytorch_model-00002-of-00002.safetensors": keys[len(keys) // 2:]}
    for name, ks in shards.items():
        save_file({k: inner[k] for k in ks}, str(folder / name))
    with open(folder / "diffusion_pytorch_model.safetensors.index.json", "w") as f:
        json.dump({"weight_map": {k: name for name, ks in shards.items() for k in ks}}, f)
    net = Wan(transformer_config=cfg, enable_logvar_linear=True)
    info = load_safetensors(net, str(folder), prefix="transformer.")
    assert info["tensors"] == len(sd) and not info["missing"] and not info["unexpected"]
    got = net.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    # a file with a tensor missing is an error under strict, reported otherwise
    one = tmp_path / "partial.safetensors"
    save_file({k: inner[k] for k in keys[1:]}, str(one))
    with pytest.raises(RuntimeError, match="missing"):
        load_safetensors(net, str(one), prefix="transformer.")
    assert load_safetensors(net, str(one), prefix="transformer.", strict=False)["missing"] == ["transformer." + keys[0]]
