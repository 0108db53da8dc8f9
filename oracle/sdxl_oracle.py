"""CPU oracle for the SDXL UNet (diffusers `UNet2DConditionModel`, SDXL-base configuration) in plain PyTorch.

TEST INFRASTRUCTURE ONLY (see oracle/edm_oracle.py): never imported by the product path.

PARITY: the reference's own SDXL code is pinned, the diffusers block internals are not.  The reference's SDXL network
(fastgen/networks/SDXL/network.py:159-416) wraps the third-party `diffusers==0.35.1` `UNet2DConditionModel` (imports at
SDXL/network.py:29-36, construction at :210-230 from the hub config `stabilityai/stable-diffusion-xl-base-1.0`), absent
from /root/reference and from this image, and the reference has no test that pins SDXL outputs.  What the
reference itself wrote runs unmodified on a stand-in UNet (oracle/ref_sdxl_harness.py): `StableDiffusionXL.forward`
(:334-416: `time_ids = [1024, 1024, 0, 0, 1024, 1024]`, pooled text embeddings, `timestep = rescale_t(t) = 1000 t`, the eps
-> x0 conversion through its own noise schedule) and the model-level `classify_forward` bound onto the UNet
(fastgen/networks/SD15/network.py:46-341: embedding sums, the skip tensors each up block receives, the output head);
tests/test_oracle_golden.py::test_sdxl_oracle_matches_reference_glue checks this file against their outputs
(tests/golden/sdxl_reference_glue.npz), bit-exact in fp32 and bf16.  Restated from the published diffusers sources, i.e.
NOT pinned: Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0), TimestepEmbedding, the "text_time" addition
embedding, ResnetBlock2D (GroupNorm 32 / eps 1e-5, SiLU, time_emb_proj added after conv1, 1x1 conv_shortcut),
Downsample2D (3x3 stride 2 pad 1), Upsample2D (nearest x2 + 3x3), Transformer2DModel with linear projections (GroupNorm
eps 1e-6), BasicTransformerBlock (LayerNorm, self-attention, cross-attention, GEGLU feed-forward) and the
`torch.cat([hidden_states, res_hidden_states])` order inside an up block.  One independent anchor for the architecture:
the parameter count of the SDXL-base configuration equals the published 2 567 463 684 (tests/test_sdxl_host_cpu.py).
State-dict keys are the diffusers ones under the reference's `unet.` prefix.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List

import torch
import torch.nn.functional as F

from .device import sdpa

SDXL_BASE = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=2,
                 down_attn=(False, True, True), transformer_layers=(1, 2, 10), num_heads=(5, 10, 20),
                 cross_attention_dim=2048, addition_time_embed_dim=256, pooled_dim=1280, norm_groups=32)


# runwayml/stable-diffusion-v1-5 unet/config.json (published; 859 520 964 parameters): attention_head_dim = 8 is read by
# diffusers' UNet2DConditionModel as the NUMBER of heads, use_linear_projection = False (1x1-conv proj_in / proj_out),
# no addition embedding
SD15_BASE = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 down_attn=(True, True, True, False), transformer_layers=(1, 1, 1, 1), num_heads=(8, 8, 8, 8),
                 cross_attention_dim=768, addition_time_embed_dim=None, pooled_dim=0, norm_groups=32,
                 use_linear_projection=False)


def tiny_sd15_config(**kw) -> Dict[str, Any]:
    cfg = dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 256, 256), layers_per_block=1,
               down_attn=(True, True, True, False), transformer_layers=(1, 1, 1, 1), num_heads=(8, 8, 8, 8),
               cross_attention_dim=96, addition_time_embed_dim=None, pooled_dim=0, norm_groups=32,
               use_linear_projection=False)
    cfg.update(kw)
    return cfg


def tiny_config(**kw) -> Dict[str, Any]:
    cfg = dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 256), layers_per_block=2,
               down_attn=(False, True, True), transformer_layers=(1, 1, 2), num_heads=(1, 2, 4),
               cross_attention_dim=64, addition_time_embed_dim=32, pooled_dim=48, norm_groups=32)
    cfg.update(kw)
    return cfg


def layout(cfg) -> Dict[str, Any]:
    """channel bookkeeping shared by the oracle, the state-dict factory and the tests"""
    boc = list(cfg["block_out_channels"])
    L = cfg["layers_per_block"]
    down = []
    cin = boc[0]
    for i, c in enumerate(boc):
        down.append(dict(res=[(cin if j == 0 else c, c) for j in range(L)], attn=cfg["down_attn"][i],
                         tl=cfg["transformer_layers"][i], heads=cfg["num_heads"][i], down=i < len(boc) - 1, c=c))
        cin = c
    rev = boc[::-1]
    up = []
    prev = rev[0]
    for i, c in enumerate(rev):
        inp = rev[min(i + 1, len(rev) - 1)]
        rs = []
        for j in range(L + 1):
            skip = inp if j == L else c
            rin = prev if j == 0 else c
            rs.append((rin, skip, c))
        up.append(dict(res=rs, attn=cfg["down_attn"][::-1][i], tl=cfg["transformer_layers"][::-1][i],
                       heads=cfg["num_heads"][::-1][i], up=i < len(rev) - 1, c=c))
        prev = c
    return dict(down=down, up=up, mid_c=boc[-1], mid_tl=cfg["transformer_layers"][-1], mid_heads=cfg["num_heads"][-1],
                temb=4 * boc[0],
                proj_in=(cfg["pooled_dim"] + 6 * cfg["addition_time_embed_dim"]) if cfg.get("addition_time_embed_dim") else 0)


def make_state_dict(cfg, seed: int) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    lay = layout(cfg)
    T, X = lay["temb"], cfg["cross_attention_dim"]
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, cout, cin, bias=True):
        sd[name + ".weight"] = torch.randn(cout, cin, generator=g) * cin ** -0.5
        if bias:
            sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (cin * k * k) ** -0.5
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def norm(name, c):
        sd[name + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    def resnet(p, cin, cout):
        norm(p + "norm1", cin)
        conv(p + "conv1", cout, cin, 3)
        lin(p + "time_emb_proj", cout, T)
        norm(p + "norm2", cout)
        conv(p + "conv2", cout, cout, 3)
        if cin != cout:
            conv(p + "conv_shortcut", cout, cin, 1)

    linear_proj = bool(cfg.get("use_linear_projection", True))

    def proj(name, c):
        if linear_proj:
            lin(name, c, c)
        else:
            conv(name, c, c, 1)

    def transformer(p, c, layers):
        norm(p + "norm", c)
        proj(p + "proj_in", c)
        for k in range(layers):
            b = f"{p}transformer_blocks.{k}."
            for nm in ("norm1", "norm2", "norm3"):
                norm(b + nm, c)
            for a, kd in (("attn1", c), ("attn2", X)):
                lin(b + a + ".to_q", c, c, bias=False)
                lin(b + a + ".to_k", c, kd, bias=False)
                lin(b + a + ".to_v", c, kd, bias=False)
                lin(b + a + ".to_out.0", c, c)
            lin(b + "ff.net.0.proj", 8 * c, c)
            lin(b + "ff.net.2", c, 4 * c)
        proj(p + "proj_out", c)

    u = "unet."
    conv(u + "conv_in", cfg["block_out_channels"][0], cfg["in_channels"], 3)
    lin(u + "time_embedding.linear_1", T, cfg["block_out_channels"][0])
    lin(u + "time_embedding.linear_2", T, T)
    if lay["proj_in"]:
        lin(u + "add_embedding.linear_1", T, lay["proj_in"])
        lin(u + "add_embedding.linear_2", T, T)
    for i, d in enumerate(lay["down"]):
        for j, (ci, co) in enumerate(d["res"]):
            resnet(f"{u}down_blocks.{i}.resnets.{j}.", ci, co)
            if d["attn"]:
                transformer(f"{u}down_blocks.{i}.attentions.{j}.", co, d["tl"])
        if d["down"]:
            conv(f"{u}down_blocks.{i}.downsamplers.0.conv", d["c"], d["c"], 3)
    mc = lay["mid_c"]
    resnet(u + "mid_block.resnets.0.", mc, mc)
    transformer(u + "mid_block.attentions.0.", mc, lay["mid_tl"])
    resnet(u + "mid_block.resnets.1.", mc, mc)
    for i, d in enumerate(lay["up"]):
        for j, (rin, skip, co) in enumerate(d["res"]):
            resnet(f"{u}up_blocks.{i}.resnets.{j}.", rin + skip, co)
            if d["attn"]:
                transformer(f"{u}up_blocks.{i}.attentions.{j}.", co, d["tl"])
        if d["up"]:
            conv(f"{u}up_blocks.{i}.upsamplers.0.conv", d["c"], d["c"], 3)
    norm(u + "conv_norm_out", cfg["block_out_channels"][0])
    conv(u + "conv_out", cfg["out_channels"], cfg["block_out_channels"][0], 3)
    return sd


# ---------------------------------------------------------------------------------------------------------
def timesteps_proj(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], fp32."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32) / half
    emb = t.cpu()[:, None].float() * torch.exp(exponent)[None, :]     # small table: always evaluated on the CPU
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1).to(t.device)


def _w(sd, name, x):
    return sd[name].to(x.dtype)


def _lin(sd, p, x):
    return F.linear(x, _w(sd, p + ".weight", x), _w(sd, p + ".bias", x) if p + ".bias" in sd else None)


def _conv(sd, p, x, stride=1, pad=1):
    return F.conv2d(x, _w(sd, p + ".weight", x), _w(sd, p + ".bias", x), stride=stride, padding=pad)


def _gn(sd, p, x, groups, eps):
    return F.group_norm(x, groups, _w(sd, p + ".weight", x), _w(sd, p + ".bias", x), eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), _w(sd, p + ".weight", x), _w(sd, p + ".bias", x), 1e-5)


def resnet(sd, p, x, emb, groups):
    h = _conv(sd, p + "conv1", F.silu(_gn(sd, p + "norm1", x, groups, 1e-5)))
    h = h + _lin(sd, p + "time_emb_proj", F.silu(emb))[:, :, None, None]
    h = _conv(sd, p + "conv2", F.silu(_gn(sd, p + "norm2", h, groups, 1e-5)))
    if p + "conv_shortcut.weight" in sd:
        x = _conv(sd, p + "conv_shortcut", x, pad=0)
    return x + h


def attention(sd, p, x, ctx, heads, bias=None):
    """diffusers Attention + AttnProcessor2_0; `bias` [B, 1, L]: the additive encoder_attention_mask of the cross-attention"""
    B, N, C = x.shape
    q = _lin(sd, p + "to_q", x).view(B, N, heads, -1).transpose(1, 2)
    k = _lin(sd, p + "to_k", ctx).view(B, ctx.shape[1], heads, -1).transpose(1, 2)
    v = _lin(sd, p + "to_v", ctx).view(B, ctx.shape[1], heads, -1).transpose(1, 2)
    if bias is None:
        o = sdpa(q, k, v)
    else:   # [B, 1, L] -> broadcast over heads and queries, added to the scaled logits
        o = sdpa(q, k, v, bias[:, None, :, :].to(q.dtype))
    o = o.transpose(1, 2).reshape(B, N, C)
    return _lin(sd, p + "to_out.0", o)


def transformer(sd, p, x, ctx, layers, heads, groups, bias=None):
    B, C, H, W = x.shape
    res = x
    linear_proj = sd[p + "proj_in.weight"].dim() == 2
    h = _gn(sd, p + "norm", x, groups, 1e-6)
    if linear_proj:
        h = _lin(sd, p + "proj_in", h.permute(0, 2, 3, 1).reshape(B, H * W, C))
    else:   # use_linear_projection = False: 1x1 convolution on the feature map, then tokens
        h = _conv(sd, p + "proj_in", h, pad=0).permute(0, 2, 3, 1).reshape(B, H * W, C)
    for k in range(layers):
        b = f"{p}transformer_blocks.{k}."
        n = _ln(sd, b + "norm1", h)
        h = attention(sd, b + "attn1.", n, n, heads) + h
        h = attention(sd, b + "attn2.", _ln(sd, b + "norm2", h), ctx, heads, bias) + h
        g = _lin(sd, b + "ff.net.0.proj", _ln(sd, b + "norm3", h))
        hid, gate = g.chunk(2, dim=-1)
        h = _lin(sd, b + "ff.net.2", hid * F.gelu(gate)) + h
    if linear_proj:
        h = _lin(sd, p + "proj_out", h).reshape(B, H, W, C).permute(0, 3, 1, 2)
    else:
        h = _conv(sd, p + "proj_out", h.reshape(B, H, W, C).permute(0, 3, 1, 2), pad=0)
    return h + res


def unet_forward(sd, cfg, sample, timestep, text, pooled, time_ids, encoder_attention_mask=None) -> torch.Tensor:
    """classify_forward (fastgen/networks/SD15/network.py:46-341) for the SDXL configuration, and -- without the addition
    embedding and with `encoder_attention_mask` [B, L] of 0 / 1 turned into the additive bias (1 - mask) * -10000 of
    :142-145 -- for the SD 1.5 configuration."""
    lay = layout(cfg)
    G = cfg["norm_groups"]
    dt = sample.dtype
    u = "unet."
    B = sample.shape[0]
    t_emb = timesteps_proj(timestep.expand(B) if timestep.ndim == 0 or timestep.numel() == 1 else timestep,
                           cfg["block_out_channels"][0]).to(dt)
    emb = _lin(sd, u + "time_embedding.linear_2", F.silu(_lin(sd, u + "time_embedding.linear_1", t_emb)))
    if cfg.get("addition_time_embed_dim"):
        te = timesteps_proj(time_ids.flatten(), cfg["addition_time_embed_dim"]).reshape(B, -1)
        add = torch.cat([pooled, te.to(pooled.dtype)], dim=-1).to(dt)
        emb = emb + _lin(sd, u + "add_embedding.linear_2", F.silu(_lin(sd, u + "add_embedding.linear_1", add)))
    ctx = text.to(dt)
    bias = None
    if encoder_attention_mask is not None:
        bias = ((1 - encoder_attention_mask.to(dt)) * -10000.0).unsqueeze(1)
    h = _conv(sd, u + "conv_in", sample)
    skips: List[torch.Tensor] = [h]
    for i, d in enumerate(lay["down"]):
        for j in range(len(d["res"])):
            h = resnet(sd, f"{u}down_blocks.{i}.resnets.{j}.", h, emb, G)
            if d["attn"]:
                h = transformer(sd, f"{u}down_blocks.{i}.attentions.{j}.", h, ctx, d["tl"], d["heads"], G, bias)
            skips.append(h)
        if d["down"]:
            h = _conv(sd, f"{u}down_blocks.{i}.downsamplers.0.conv", h, stride=2, pad=1)
            skips.append(h)
    h = resnet(sd, u + "mid_block.resnets.0.", h, emb, G)
    h = transformer(sd, u + "mid_block.attentions.0.", h, ctx, lay["mid_tl"], lay["mid_heads"], G, bias)
    h = resnet(sd, u + "mid_block.resnets.1.", h, emb, G)
    for i, d in enumerate(lay["up"]):
        for j in range(len(d["res"])):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(sd, f"{u}up_blocks.{i}.resnets.{j}.", h, emb, G)
            if d["attn"]:
                h = transformer(sd, f"{u}up_blocks.{i}.attentions.{j}.", h, ctx, d["tl"], d["heads"], G, bias)
        if d["up"]:
            h = _conv(sd, f"{u}up_blocks.{i}.upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    h = F.silu(_gn(sd, u + "conv_norm_out", h, G, 1e-5))
    return _conv(sd, u + "conv_out", h)


def alphas_cumprod() -> torch.Tensor:
    """DDIM scheduler of SDXL-base: scaled_linear betas 0.00085 -> 0.012, 1000 steps."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def sdxl_forward(sd, cfg, x_t, t, text, pooled, fwd_pred_type="x0", dtype=torch.float32, resolution=1024):
    """StableDiffusionXL.forward (SDXL/network.py:334-416): eps prediction, 'sdxl' (alphas) schedule."""
    B = x_t.shape[0]
    time_ids = torch.tensor([[resolution, resolution, 0, 0, resolution, resolution]], dtype=torch.float32,
                            device=x_t.device).repeat(B, 1)
    eps = unet_forward(sd, cfg, x_t.to(dtype), (t.double() * 1000.0).to(t.dtype), text, pooled.to(dtype), time_ids)
    eps = eps.to(x_t.dtype)
    if fwd_pred_type == "eps":
        return eps
    abar = alphas_cumprod().double().to(x_t.device)
    idx = (1000 * t.double()).long().clamp(0, 999)
    a, s = abar[idx].sqrt().reshape(-1, 1, 1, 1), (1 - abar[idx]).sqrt().reshape(-1, 1, 1, 1)
    a = torch.where(a >= 0, a.clamp(min=1e-6), a.clamp(max=-1e-6))
    return ((x_t.double() - eps.double() * s) / a).to(x_t.dtype)


def sd15_forward(sd, cfg, x_t, t, text, mask=None, fwd_pred_type="x0", dtype=torch.float32):
    """StableDiffusion15.forward (SD15/network.py:602-678): eps prediction, 'sd' (alphas) schedule,
    condition = [encoder_hidden_states, encoder_attention_mask]."""
    eps = unet_forward(sd, cfg, x_t.to(dtype), (t.double() * 1000.0).to(t.dtype), text, None, None,
                       encoder_attention_mask=mask)
    eps = eps.to(x_t.dtype)
    if fwd_pred_type == "eps":
        return eps
    abar = alphas_cumprod().double().to(x_t.device)
    idx = (1000 * t.double()).long().clamp(0, 999)
    a, s = abar[idx].sqrt().reshape(-1, 1, 1, 1), (1 - abar[idx]).sqrt().reshape(-1, 1, 1, 1)
    a = torch.where(a >= 0, a.clamp(min=1e-6), a.clamp(max=-1e-6))
    return ((x_t.double() - eps.double() * s) / a).to(x_t.dtype)
