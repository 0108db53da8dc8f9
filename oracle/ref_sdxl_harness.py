"""Runs the UNMODIFIED reference SDXL code on CPU against a stand-in diffusers UNet.

TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).  The reference's `StableDiffusionXL.forward`
(fastgen/networks/SDXL/network.py:334-416) and the model-level `classify_forward` it binds onto the UNet
(fastgen/networks/SD15/network.py:46-341) run as they are; `UNet2DConditionModel` and its blocks are stand-ins whose
arithmetic is oracle/sdxl_oracle.py's restatement of the published diffusers modules (ResnetBlock2D, Transformer2DModel,
Down/Upsample2D, the text_time addition embedding).  Pinned that way: the reference's conditioning inputs (time ids,
pooled embeddings, 1000 t), the order in which the embeddings are summed, which skip tensors each up block receives
(`down_block_res_samples[-len(upsample_block.resnets):]`), the output head and the eps -> x0 conversion through the
reference's own noise schedule.  NOT pinned (inside diffusers): the blocks themselves, incl. the order of
`torch.cat([hidden_states, res_hidden_states])` inside an up block.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Any, Dict

import torch
import torch.nn.functional as F
from torch import nn

from oracle import sdxl_oracle as XO
from oracle.ref_shims import REF_ROOT, install_shims


class _Fn(nn.Module):
    """a module whose forward is a closure over the oracle's state dict (the stand-ins carry no parameters of their own)"""

    def __init__(self, fn):
        super().__init__()
        self._fn = fn

    def forward(self, *a, **kw):
        return self._fn(*a, **kw)


class _DownBlock(nn.Module):
    def __init__(self, sd, cfg, i, d, groups):
        super().__init__()
        self.sd, self.i, self.d, self.groups = sd, i, d, groups
        self.has_cross_attention = bool(d["attn"])
        self.resnets = nn.ModuleList([nn.Identity() for _ in d["res"]])

    def forward(self, hidden_states, temb, encoder_hidden_states=None, **kw):
        u, out = "unet.", ()
        for j in range(len(self.d["res"])):
            hidden_states = XO.resnet(self.sd, f"{u}down_blocks.{self.i}.resnets.{j}.", hidden_states, temb, self.groups)
            if self.d["attn"]:
                hidden_states = XO.transformer(self.sd, f"{u}down_blocks.{self.i}.attentions.{j}.", hidden_states,
                                               encoder_hidden_states, self.d["tl"], self.d["heads"], self.groups,
                                               kw.get("encoder_attention_mask"))
            out += (hidden_states,)
        if self.d["down"]:
            hidden_states = XO._conv(self.sd, f"{u}down_blocks.{self.i}.downsamplers.0.conv", hidden_states, stride=2, pad=1)
            out += (hidden_states,)
        return hidden_states, out


class _MidBlock(nn.Module):
    has_cross_attention = True

    def __init__(self, sd, lay, groups):
        super().__init__()
        self.sd, self.lay, self.groups = sd, lay, groups

    def forward(self, hidden_states, temb, encoder_hidden_states=None, **kw):
        u = "unet."
        h = XO.resnet(self.sd, u + "mid_block.resnets.0.", hidden_states, temb, self.groups)
        h = XO.transformer(self.sd, u + "mid_block.attentions.0.", h, encoder_hidden_states, self.lay["mid_tl"],
                           self.lay["mid_heads"], self.groups, kw.get("encoder_attention_mask"))
        return XO.resnet(self.sd, u + "mid_block.resnets.1.", h, temb, self.groups)


class _UpBlock(nn.Module):
    def __init__(self, sd, i, d, groups):
        super().__init__()
        self.sd, self.i, self.d, self.groups = sd, i, d, groups
        self.has_cross_attention = bool(d["attn"])
        self.resnets = nn.ModuleList([nn.Identity() for _ in d["res"]])

    def forward(self, hidden_states, temb, res_hidden_states_tuple, encoder_hidden_states=None, upsample_size=None, **kw):
        u = "unet."
        for j in range(len(self.d["res"])):
            res = res_hidden_states_tuple[-1]                       # diffusers: pop from the end, cat [hidden, res]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = XO.resnet(self.sd, f"{u}up_blocks.{self.i}.resnets.{j}.", hidden_states, temb, self.groups)
            if self.d["attn"]:
                hidden_states = XO.transformer(self.sd, f"{u}up_blocks.{self.i}.attentions.{j}.", hidden_states,
                                               encoder_hidden_states, self.d["tl"], self.d["heads"], self.groups,
                                               kw.get("encoder_attention_mask"))
        if self.d["up"]:
            hidden_states = XO._conv(self.sd, f"{u}up_blocks.{self.i}.upsamplers.0.conv",
                                     F.interpolate(hidden_states, scale_factor=2.0, mode="nearest"))
        return hidden_states


class UNet2DConditionModel(nn.Module):
    """stand-in with the attributes / helper methods `classify_forward` uses"""

    def __init__(self, cfg: Dict[str, Any], sd: Dict[str, torch.Tensor]):
        super().__init__()
        lay, G, u = XO.layout(cfg), cfg["norm_groups"], "unet."
        self.sd, self.cfg = sd, cfg
        self.config = types.SimpleNamespace(center_input_sample=False, class_embeddings_concat=False,
                                            addition_embed_type="text_time" if cfg.get("addition_time_embed_dim") else None)
        self.num_upsamplers = len(cfg["block_out_channels"]) - 1
        self.time_embed_act = None
        self.time_embedding = _Fn(lambda t_emb, cond=None: XO._lin(sd, u + "time_embedding.linear_2",
                                                                   F.silu(XO._lin(sd, u + "time_embedding.linear_1", t_emb))))
        self.conv_in = _Fn(lambda x: XO._conv(sd, u + "conv_in", x))
        self.down_blocks = nn.ModuleList([_DownBlock(sd, cfg, i, d, G) for i, d in enumerate(lay["down"])])
        self.mid_block = _MidBlock(sd, lay, G)
        self.up_blocks = nn.ModuleList([_UpBlock(sd, i, d, G) for i, d in enumerate(lay["up"])])
        self.conv_norm_out = _Fn(lambda x: XO._gn(sd, u + "conv_norm_out", x, G, 1e-5))
        self.conv_act = nn.SiLU()
        self.conv_out = _Fn(lambda x: XO._conv(sd, u + "conv_out", x))

    def get_time_embed(self, sample, timestep):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.float64 if isinstance(timestep, float) else torch.int64)
        elif timesteps.ndim == 0:
            timesteps = timesteps[None]
        timesteps = timesteps.expand(sample.shape[0])
        return XO.timesteps_proj(timesteps, self.cfg["block_out_channels"][0]).to(dtype=sample.dtype)

    def get_class_embed(self, sample, class_labels):
        return None

    def enable_gradient_checkpointing(self):      # StableDiffusion15.forward calls it on every forward (SD15/network.py:643)
        pass

    def get_aug_embed(self, emb, encoder_hidden_states, added_cond_kwargs):
        if not self.cfg.get("addition_time_embed_dim"):   # diffusers: addition_embed_type None -> no additional embedding
            return None
        text_embeds, time_ids = added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"]
        te = XO.timesteps_proj(time_ids.flatten(), self.cfg["addition_time_embed_dim"]).reshape(text_embeds.shape[0], -1)
        add = torch.cat([text_embeds, te], dim=-1).to(emb.dtype)
        u = "unet."
        return XO._lin(self.sd, u + "add_embedding.linear_2", F.silu(XO._lin(self.sd, u + "add_embedding.linear_1", add)))

    def process_encoder_hidden_states(self, encoder_hidden_states, added_cond_kwargs):
        return encoder_hidden_states


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_sdxl():
    if not os.path.isdir(os.path.join(REF_ROOT, "fastgen")):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    install_shims()
    d = sys.modules.get("diffusers") or _mod("diffusers")
    d.UNet2DConditionModel, d.AutoencoderKL = UNet2DConditionModel, type("AutoencoderKL", (), {})
    blk = {n: type(n, (nn.Module,), {}) for n in ("CrossAttnDownBlock2D", "CrossAttnUpBlock2D", "DownBlock2D", "UpBlock2D",
                                                  "UNetMidBlock2DCrossAttn")}
    dm = sys.modules.get("diffusers.models") or _mod("diffusers.models")
    du_ = _mod("diffusers.models.unets")
    db = _mod("diffusers.models.unets.unet_2d_blocks", **blk)
    utils = sys.modules.get("diffusers.utils") or _mod("diffusers.utils")
    utils.USE_PEFT_BACKEND = False
    utils.unscale_lora_layers = lambda *a, **k: None
    utils.scale_lora_layers = lambda *a, **k: None
    utils.deprecate = lambda *a, **k: None
    d.models, d.utils, dm.unets, du_.unet_2d_blocks = dm, utils, du_, db
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("fastgen.networks.SDXL.network")


def reference_sdxl_forward(cfg, sd, x_t, t, text, pooled, fwd_pred_type="x0"):
    """`StableDiffusionXL.forward` (unbound, on a minimal stand-in for `self`) driving `classify_forward` on the stand-in UNet;
    the noise schedule is the reference's own AlphasNoiseSchedule over the SDXL scheduler's alphas_cumprod"""
    m = load_reference_sdxl()
    from fastgen.networks.noise_schedule import get_noise_schedule

    unet = UNet2DConditionModel(cfg, sd)
    unet.forward = types.MethodType(m.classify_forward, unet)
    me = types.SimpleNamespace(unet=unet, net_pred_type="eps",
                               noise_scheduler=get_noise_schedule("alphas", alphas_cumprod=XO.alphas_cumprod()),
                               add_time_ids=m.StableDiffusionXL.build_condition_input(None, 1024))
    return m.StableDiffusionXL.forward(me, x_t, t, condition=[text, pooled], fwd_pred_type=fwd_pred_type)


def reference_sd15_forward(cfg, sd, x_t, t, text, mask, fwd_pred_type="x0"):
    """`StableDiffusion15.forward` (SD15/network.py:602-678, unbound, on a minimal stand-in for `self`) driving
    `classify_forward` with `encoder_attention_mask` on the stand-in UNet at an SD 1.5-shaped configuration (no addition
    embedding, 1x1-conv projections, four levels); the noise schedule is the reference's AlphasNoiseSchedule over the SD
    scheduler's alphas_cumprod (the same scaled-linear betas as SDXL)"""
    load_reference_sdxl()
    m15 = importlib.import_module("fastgen.networks.SD15.network")
    from fastgen.networks.noise_schedule import get_noise_schedule

    unet = UNet2DConditionModel(cfg, sd)
    unet.forward = types.MethodType(m15.classify_forward, unet)
    me = types.SimpleNamespace(unet=unet, net_pred_type="eps",
                               noise_scheduler=get_noise_schedule("alphas", alphas_cumprod=XO.alphas_cumprod()))
    return m15.StableDiffusion15.forward(me, x_t, t, condition=[text, mask], fwd_pred_type=fwd_pred_type)
