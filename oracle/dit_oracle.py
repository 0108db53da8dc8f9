"""CPU oracle for the DiT network (adaLN-zero transformer) in plain PyTorch.

TEST INFRASTRUCTURE ONLY (see oracle/edm_oracle.py): never imported by the product path.

PARITY: PINNED to the reference's own DiT code; only timm's three classes are restated.  The reference's DiT
(fastgen/networks/DiT/network.py) builds its blocks from the third-party `timm` package
(`timm.models.vision_transformer.{PatchEmbed, Attention, Mlp}`, imported at DiT/network.py:15-16; un-pinned in
requirements.txt:20), which is neither under /root/reference nor installed here, and the reference's own test
(tests/test_network.py:362-410) asserts shapes only.  The UNMODIFIED reference `DiT` class runs with stand-ins for just those
three classes (oracle/ref_dit_harness.py) and tests/test_oracle_golden.py::test_dit_oracle_matches_reference checks this
file against its outputs (tests/golden/dit_reference.npz: plain, 72-wide heads, SiT convention, r embedder diff / abs, x0
and flow, the position table) -- bit-exact.  Followed line by line from the reference:

  DiT/network.py:29-41     apply_adaptive_modulation
  DiT/network.py:49-101    FourierTimeEmbedding ([cos | sin], exp(-ln(1e4) i / half), proj_net Linear-SiLU-Linear)
  DiT/network.py:104-145   ConditionalEmbedding (class table with one extra null row)
  DiT/network.py:153-201   DiTBlock (chunk order attn_shift, attn_scale, attn_gate, ffn_shift, ffn_scale, ffn_gate)
  DiT/network.py:204-225   OutputProjection
  DiT/network.py:437-462   unpatchify ("bhwpqc->bchpwq"), prepare_t (rescale_t = 1000 t for the RF schedule)
  DiT/network.py:464-574   forward (one-hot -> index with the all-zero row -> null class, SiT convention flips)
  DiT/network.py:712-790   compute_sinusoidal_2d_embeddings ([sin | cos] per axis, x axis first)

and restated from the published timm sources: PatchEmbed = Conv2d(kernel = stride = patch) + flatten(2).transpose(1, 2);
Attention(qkv_bias=True, no qk norm) = Linear(D, 3D) viewed [B, N, 3, heads, d], softmax(q k^T / sqrt(d)) v, Linear(D, D);
Mlp = Linear -> GELU(tanh) -> Linear.
"""
from __future__ import annotations

import math
from typing import Any, Dict

import numpy as np
import torch
import torch.nn.functional as F


def tiny_config(**kw) -> Dict[str, Any]:
    cfg = dict(input_size=8, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, mlp_ratio=4.0,
               num_classes=10, class_dropout_prob=0.1, scale_t=True, use_sit_convention=False)
    cfg.update(kw)
    return cfg


DIT_XL = dict(input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16, mlp_ratio=4.0,
              num_classes=1000, class_dropout_prob=0.1, scale_t=True, use_sit_convention=False)


def _encode_1d(positions: np.ndarray, dim: int, max_wavelength: float = 10000.0) -> np.ndarray:
    half = dim // 2
    freq = 1.0 / (max_wavelength ** (np.arange(half, dtype=np.float64) / float(half)))
    ang = np.outer(positions, freq)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def pos_embed_2d(dim: int, size: int) -> torch.Tensor:
    y, x = np.meshgrid(np.arange(size, dtype=np.float32), np.arange(size, dtype=np.float32), indexing="ij")
    emb = np.concatenate([_encode_1d(x.reshape(-1), dim // 2), _encode_1d(y.reshape(-1), dim // 2)], axis=1)
    return torch.from_numpy(emb).float().unsqueeze(0)


def make_state_dict(cfg: Dict[str, Any], seed: int) -> Dict[str, torch.Tensor]:
    """Random non-degenerate weights under the reference's key names (the zero-initialised layers are NOT zero)."""
    g = torch.Generator().manual_seed(seed)
    D, C, p = cfg["hidden_size"], cfg["in_channels"], cfg["patch_size"]
    H = int(D * cfg["mlp_ratio"])
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, cout, cin, std=None):
        sd[name + ".weight"] = torch.randn(cout, cin, generator=g) * (std if std is not None else cin ** -0.5)
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    sd["x_embedder.proj.weight"] = torch.randn(D, C, p, p, generator=g) * (C * p * p) ** -0.5
    sd["x_embedder.proj.bias"] = torch.randn(D, generator=g) * 0.02
    lin("t_embedder.proj_net.0", D, 256)
    lin("t_embedder.proj_net.2", D, D)
    n_emb = cfg["num_classes"] + (1 if cfg["class_dropout_prob"] > 0 else 0)
    sd["y_embedder.class_embeddings.weight"] = torch.randn(n_emb, D, generator=g) * 0.5
    sd["pos_embed"] = pos_embed_2d(D, cfg["input_size"] // p)
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        lin(b + "attention.qkv", 3 * D, D)
        lin(b + "attention.proj", D, D)
        lin(b + "feed_forward.fc1", H, D)
        lin(b + "feed_forward.fc2", D, H)
        lin(b + "conditioning_net.1", 6 * D, D, std=0.5 * D ** -0.5)
    lin("final_layer.adaptive_params.1", 2 * D, D, std=0.5 * D ** -0.5)
    lin("final_layer.projection", p * p * C, D)
    lin("logvar_linear", 1, D)
    if cfg.get("r_timestep"):     # drawn last: the other weights do not depend on this option
        lin("r_embedder.proj_net.0", D, 256)
        lin("r_embedder.proj_net.2", D, D)
    return sd


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype))


def _modulate(x, shift, scale):
    return x * (1.0 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def dit_forward(sd, cfg, x_t: torch.Tensor, t: torch.Tensor, condition: torch.Tensor, fwd_pred_type: str = "x0",
                net_pred_type: str = "flow", r: torch.Tensor = None, time_cond_type: str = "abs") -> torch.Tensor:
    """DiT.forward (DiT/network.py:464-574) with the RF schedule; dtype of x_t is the compute dtype."""
    D, p, heads = cfg["hidden_size"], cfg["patch_size"], cfg["num_heads"]
    x_in, t_in = x_t, t
    if condition.ndim == 2:
        mask = torch.any(condition != 0, dim=1)
        condition = torch.where(~mask, cfg["num_classes"], condition.argmax(dim=1))
    tt = (t.double() * 1000.0).to(t.dtype) if cfg["scale_t"] else t     # rescale_t (noise_schedule.py:1337-1341)
    tt = tt.to(x_t.dtype)
    if cfg["use_sit_convention"]:
        tt = 1 - tt
    x = F.conv2d(x_t, sd["x_embedder.proj.weight"].to(x_t.dtype), sd["x_embedder.proj.bias"].to(x_t.dtype), stride=p)
    x = x.flatten(2).transpose(1, 2) + sd["pos_embed"].to(x_t.dtype)
    half = 128
    freq = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)

    def fourier(v):
        ang = v[:, None].float() * freq[None, :]
        return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(v.dtype)

    t_emb = _lin(sd, "t_embedder.proj_net.2", F.silu(_lin(sd, "t_embedder.proj_net.0", fourier(tt))))
    y = sd["y_embedder.class_embeddings.weight"].to(x_t.dtype)[condition]
    c = t_emb + y
    if r is not None and "r_embedder.proj_net.0.weight" in sd:     # DiT/network.py:518-533 (r is NOT flipped for SiT)
        rr = ((r.double() * 1000.0).to(r.dtype) if cfg["scale_t"] else r).to(x_t.dtype)
        rin = tt - rr if time_cond_type == "diff" else rr
        c = c + _lin(sd, "r_embedder.proj_net.2", F.silu(_lin(sd, "r_embedder.proj_net.0", fourier(rin))))
    B, N, _ = x.shape
    d = D // heads
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        a_shift, a_scale, a_gate, f_shift, f_scale, f_gate = _lin(sd, b + "conditioning_net.1", F.silu(c)).chunk(6, dim=1)
        h = _modulate(F.layer_norm(x, (D,), eps=1e-6), a_shift, a_scale)
        qkv = _lin(sd, b + "attention.qkv", h).reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q * d ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, N, D)
        x = x + a_gate.unsqueeze(1) * _lin(sd, b + "attention.proj", o)
        h = _modulate(F.layer_norm(x, (D,), eps=1e-6), f_shift, f_scale)
        ff = _lin(sd, b + "feed_forward.fc2", F.gelu(_lin(sd, b + "feed_forward.fc1", h), approximate="tanh"))
        x = x + f_gate.unsqueeze(1) * ff
    shift, scale = _lin(sd, "final_layer.adaptive_params.1", F.silu(c)).chunk(2, dim=1)
    x = _lin(sd, "final_layer.projection", _modulate(F.layer_norm(x, (D,), eps=1e-6), shift, scale))
    hh = int(N ** 0.5)
    cc = cfg["in_channels"]
    x = torch.einsum("bhwpqc->bchpwq", x.reshape(B, hh, hh, p, p, cc)).reshape(B, cc, hh * p, hh * p)
    if cfg["use_sit_convention"] and net_pred_type == "flow":
        x = -x
    if fwd_pred_type == net_pred_type:
        return x
    assert net_pred_type == "flow" and fwd_pred_type == "x0"
    return (x_in.double() - t_in.double().reshape(-1, 1, 1, 1) * x.double()).to(x_in.dtype)
