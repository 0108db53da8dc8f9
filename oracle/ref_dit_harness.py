"""Runs the UNMODIFIED reference DiT (fastgen/networks/DiT/network.py) on CPU with stand-ins for the three timm classes it
imports (`PatchEmbed`, `Attention`, `Mlp`: timm is not installed here).

TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).  Everything else -- DiT.__init__ / initialize_weights
/ forward, DiTBlock, OutputProjection, FourierTimeEmbedding, ClassEmbedding, the 2-D sinusoidal position table, the
adaptive modulation, unpatchify, the SiT convention, the r embedder and the conversion to the requested prediction type --
is the reference's code.  The stand-ins restate timm's published definitions: PatchEmbed = Conv2d(patch, stride patch) +
flatten(2).transpose(1, 2); Attention (fused path off, as the reference sets it: DiT/network.py:251-255) = qkv Linear ->
[B, H, N, d] -> softmax(q * d^-0.5 @ k^T) @ v -> proj; Mlp = fc1 -> act -> fc2.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

from torch import nn

from oracle.ref_shims import REF_ROOT, install_shims


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, bias=True):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, patch_size, patch_size, bias=bias)
        self.num_patches = (img_size // patch_size) ** 2
        self.patch_size = (patch_size, patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, **kwargs):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q * self.scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(in_features, hidden_features), act_layer(), nn.Linear(hidden_features, in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_dit():
    if not os.path.isdir(os.path.join(REF_ROOT, "fastgen")):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    install_shims()
    if "timm" not in sys.modules:
        t = _mod("timm")
        t.layers = _mod("timm.layers", set_fused_attn=lambda *a, **k: None)
        t.models = _mod("timm.models")
        t.models.vision_transformer = _mod("timm.models.vision_transformer", PatchEmbed=PatchEmbed, Attention=Attention, Mlp=Mlp)
    d = sys.modules.get("diffusers") or _mod("diffusers")
    if not hasattr(d, "DDIMScheduler"):
        d.DDIMScheduler = type("DDIMScheduler", (), {})
    dm = sys.modules.get("diffusers.models") or _mod("diffusers.models")
    dm.AutoencoderKL = getattr(dm, "AutoencoderKL", type("AutoencoderKL", (), {}))
    d.models = dm
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("fastgen.networks.DiT.network")


def build_reference_dit(cfg, sd):
    """the reference `DiT(**cfg)` carrying the oracle's weights"""
    m = load_reference_dit()
    net = m.DiT(**cfg)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [k for k in missing if "pos_embed" not in k], missing
    return net.eval()
