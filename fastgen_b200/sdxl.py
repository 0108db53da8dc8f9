"""Stable Diffusion XL UNet (eps-prediction student of the DMD2 4-step sampler) on the B200 kernel library.

Drop-in for ``fastgen.networks.SDXL.network.StableDiffusionXL`` (reference SDXL/network.py:159-416): same constructor
keywords, ``forward(x_t, t, condition=[prompt_embeds, pooled], ..., fwd_pred_type)`` signature, ``add_time_ids`` /
``build_condition_input`` and the diffusers ``UNet2DConditionModel`` ``state_dict`` key names under ``unet.`` (the
reference builds the UNet from the hub config ``stabilityai/stable-diffusion-xl-base-1.0``; that configuration is
restated in :data:`SDXL_BASE` -- SURVEY.md A.4 -- so no hub access is needed).  Inference only: ``r``, feature and
logvar outputs raise like the reference does for ``r`` (SDXL/network.py:365-367) or are training-time paths.

Forward (model-level ``classify_forward``, fastgen/networks/SD15/network.py:46-341) as a flat launch list:

  fg_edm_input (NCHW -> channels-last bf16) -> conv_in
  ResnetBlock2D   = fg_gn_apply(norm1+SiLU [+ raw copy for the 1x1 shortcut]) -> fg_conv_gemm(conv1 + bias + time
                    projection row vector, GroupNorm statistics of norm2 fused) -> fg_gn_apply(norm2+SiLU) ->
                    fg_conv_gemm(conv2 [+ 1x1 shortcut as a second K segment | + identity residual], statistics of the
                    next norm fused)
  Transformer2D   = fg_gn_apply(norm) -> GEMM(proj_in) -> per block [fg_ln_modulate -> GEMM(q|k|v) -> fg_flash_attn ->
                    GEMM(to_out + residual) -> fg_ln_modulate -> GEMM(q) -> fg_flash_attn(77 text keys) -> GEMM(to_out +
                    residual) -> fg_ln_modulate -> GEMM(ff.0 with the GEGLU gate in its tail) -> GEMM(ff.2 + residual)] ->
                    GEMM(proj_out + the block input as residual)
  Downsample2D    = fg_cast_resample -> fg_conv_gemm (stride-1 taps) -> fg_strided_copy (keeps the even pixels: the
                    stride-2 / pad-1 convolution evaluated where its outputs are)
  Upsample2D      = fg_cast_resample(nearest 2x) -> fg_conv_gemm
  skip concatenations are zero-copy: the producer writes its slice of the consumer's channels-last buffer.

The feature maps between layers are fp32 (GroupNorm statistics are accumulated by the producing GEMM), every tensor-core
operand is bf16, the token (residual) stream inside the transformer blocks is fp32 like the feature maps, accumulation /
statistics / softmax are fp32.  The
text keys/values of all 70 cross-attention layers depend on the prompt only and are computed once per condition tensor.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, List, Mapping, Optional, Sequence, Set, Tuple

import torch
from torch import nn

from . import _lib as L
from . import ops
from .network import FastGenNetwork, tensor_identity
from .noise_schedule import NET_PRED_TYPES
from .wan import _Plan

__all__ = ["StableDiffusionXL", "StableDiffusion15", "SDXL_BASE", "SD15_BASE"]

# tests/test_sdxl_host_cpu.py sets this to build the launch list against host memory, so that the argument validation
# of every C-ABI call can be exercised on a CPU-only host; such a plan cannot execute (every launch fails).
_ALLOW_HOST_PLANS = False

# stabilityai/stable-diffusion-xl-base-1.0, unet/config.json
SDXL_BASE: Dict[str, Any] = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=2,
    down_attn=(False, True, True), transformer_layers=(1, 2, 10), num_heads=(5, 10, 20),
    cross_attention_dim=2048, addition_time_embed_dim=256, pooled_dim=1280, norm_groups=32)

# runwayml/stable-diffusion-v1-5, unet/config.json: four levels, attention on the first three, ONE transformer layer each,
# `attention_head_dim = 8` read by diffusers as 8 HEADS per level (head widths 40 / 80 / 160 / 160), 768-wide CLIP text
# states, 1x1-convolution proj_in / proj_out (`use_linear_projection = False`), no additional embedding.
SD15_BASE: Dict[str, Any] = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    down_attn=(True, True, True, False), transformer_layers=(1, 1, 1, 1), num_heads=(8, 8, 8, 8),
    cross_attention_dim=768, addition_time_embed_dim=None, pooled_dim=0, norm_groups=32, use_linear_projection=False)


def layout(cfg: Mapping[str, Any]) -> Dict[str, Any]:
    """channel bookkeeping of UNet2DConditionModel's down / mid / up blocks for a configuration"""
    boc = list(cfg["block_out_channels"])
    nl = cfg["layers_per_block"]
    down, cin = [], boc[0]
    for i, c in enumerate(boc):
        down.append(dict(res=[(cin if j == 0 else c, c) for j in range(nl)], attn=cfg["down_attn"][i],
                         tl=cfg["transformer_layers"][i], heads=cfg["num_heads"][i], down=i < len(boc) - 1, c=c))
        cin = c
    rev = boc[::-1]
    up, prev = [], rev[0]
    for i, c in enumerate(rev):
        inp = rev[min(i + 1, len(rev) - 1)]
        rs = [((prev if j == 0 else c), (inp if j == nl else c), c) for j in range(nl + 1)]
        up.append(dict(res=rs, attn=cfg["down_attn"][::-1][i], tl=cfg["transformer_layers"][::-1][i],
                       heads=cfg["num_heads"][::-1][i], up=i < len(rev) - 1, c=c))
        prev = c
    return dict(down=down, up=up, mid_c=boc[-1], mid_tl=cfg["transformer_layers"][-1], mid_heads=cfg["num_heads"][-1],
                temb=4 * boc[0],
                proj_in=(cfg["pooled_dim"] + 6 * cfg["addition_time_embed_dim"]) if cfg.get("addition_time_embed_dim") else 0)


# ---------------------------------------------------------------------------------------------------------
# parameter containers with the diffusers module / key names (no forward: the launch list below is the forward)
# ---------------------------------------------------------------------------------------------------------
class _TimestepEmbedding(nn.Module):
    def __init__(self, cin: int, dim: int):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(cin, dim), nn.Linear(dim, dim)


class _Resnet(nn.Module):
    def __init__(self, cin: int, cout: int, temb: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)
        self.cin, self.cout = cin, cout


class _Attention(nn.Module):
    def __init__(self, dim: int, kv_dim: int):
        super().__init__()
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(kv_dim, dim, bias=False)
        self.to_v = nn.Linear(kv_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])


class _GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, 2 * inner)


class _FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(dim, 4 * dim), nn.Identity(), nn.Linear(4 * dim, dim)])


class _BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, cross_dim: int):
        super().__init__()
        self.norm1, self.attn1 = nn.LayerNorm(dim), _Attention(dim, dim)
        self.norm2, self.attn2 = nn.LayerNorm(dim), _Attention(dim, cross_dim)
        self.norm3, self.ff = nn.LayerNorm(dim), _FeedForward(dim)


class _Transformer2D(nn.Module):
    def __init__(self, dim: int, layers: int, heads: int, cross_dim: int, groups: int, linear: bool = True):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        # use_linear_projection: SDXL projects tokens with nn.Linear, SD 1.5 the feature map with a 1x1 convolution --
        # the same arithmetic, different parameter shapes ([C, C] vs [C, C, 1, 1]) in the state dict
        self.proj_in = nn.Linear(dim, dim) if linear else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([_BasicTransformerBlock(dim, cross_dim) for _ in range(layers)])
        self.proj_out = nn.Linear(dim, dim) if linear else nn.Conv2d(dim, dim, 1)
        self.dim, self.heads = dim, heads


class _Sampler2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)


class _UNetBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList()


class _UNet(nn.Module):
    def __init__(self, cfg: Mapping[str, Any]):
        super().__init__()
        lay = layout(cfg)
        boc, G, T, X = cfg["block_out_channels"], cfg["norm_groups"], lay["temb"], cfg["cross_attention_dim"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = _TimestepEmbedding(boc[0], T)
        if lay["proj_in"]:
            self.add_embedding = _TimestepEmbedding(lay["proj_in"], T)
        lin = bool(cfg.get("use_linear_projection", True))
        self.down_blocks, self.up_blocks = nn.ModuleList(), nn.ModuleList()
        for d in lay["down"]:
            blk = _UNetBlock()
            blk.resnets.extend(_Resnet(ci, co, T, G) for ci, co in d["res"])
            if d["attn"]:
                blk.attentions = nn.ModuleList(_Transformer2D(d["c"], d["tl"], d["heads"], X, G, lin) for _ in d["res"])
            if d["down"]:
                blk.downsamplers = nn.ModuleList([_Sampler2D(d["c"])])
            self.down_blocks.append(blk)
        mc = lay["mid_c"]
        self.mid_block = _UNetBlock()
        self.mid_block.resnets.extend([_Resnet(mc, mc, T, G), _Resnet(mc, mc, T, G)])
        self.mid_block.attentions = nn.ModuleList([_Transformer2D(mc, lay["mid_tl"], lay["mid_heads"], X, G, lin)])
        for u in lay["up"]:
            blk = _UNetBlock()
            blk.resnets.extend(_Resnet(rin + skip, co, T, G) for rin, skip, co in u["res"])
            if u["attn"]:
                blk.attentions = nn.ModuleList(_Transformer2D(u["c"], u["tl"], u["heads"], X, G, lin) for _ in u["res"])
            if u["up"]:
                blk.upsamplers = nn.ModuleList([_Sampler2D(u["c"])])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(G, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)
        self.logvar_linear = nn.Linear(boc[-1], 1)   # reference SDXL/network.py:231 (training only)


class _Act:
    """channels [off, off+c) of a channels-last fp32 buffer [n, h, w, ld]"""
    __slots__ = ("buf", "off", "c")

    def __init__(self, buf: torch.Tensor, off: int, c: int):
        self.buf, self.off, self.c = buf, off, c

    @property
    def h(self) -> int:
        return self.buf.shape[1]

    @property
    def w(self) -> int:
        return self.buf.shape[2]

    @property
    def ld(self) -> int:
        return self.buf.shape[3]

    @property
    def t(self) -> torch.Tensor:
        return self.buf[..., self.off:self.off + self.c]


def timestep_sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)``: [cos | sin] in fp32"""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    ang = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


class StableDiffusionXL(FastGenNetwork):
    MODEL_ID = "stabilityai/stable-diffusion-xl-base-1.0"

    def __init__(self, model_id: str = MODEL_ID, net_pred_type: str = "eps", schedule_type: str = "sdxl",
                 load_pretrained: bool = True, unet_config: Optional[Mapping[str, Any]] = None, **model_kwargs):
        super().__init__(net_pred_type=net_pred_type, schedule_type=schedule_type, **model_kwargs)
        self.model_id = model_id
        self.cfg = dict(SDXL_BASE if unet_config is None else unet_config)
        for c in self.cfg["block_out_channels"]:
            if c % self.cfg["norm_groups"] or c % 8:
                raise ValueError("block_out_channels must be multiples of 8 and of norm_groups")
        for c, hds in zip(self.cfg["block_out_channels"], self.cfg["num_heads"]):
            if c % hds or (c // hds) % 8 or c // hds > 256:
                raise ValueError("attention head dim must be a multiple of 8 and <= 256")
        # there is no hub here: `load_pretrained` is accepted for signature compatibility, weights arrive through
        # load_state_dict() with the diffusers key names (reference utils/checkpointer.py:88-108)
        self.unet = _UNet(self.cfg)
        self.add_time_ids = self.build_condition_input(1024)
        self._has_add = bool(self.cfg.get("addition_time_embed_dim"))
        self.reset_parameters()
        self._packed: Optional[Dict[str, Any]] = None
        self._plans: Dict[Any, Any] = {}
        self._ctx: Optional[Dict[str, Any]] = None
        self.profile: Optional[Dict[str, Dict[str, float]]] = None

    # ---- reference API ----------------------------------------------------------------------------------
    def reset_parameters(self):
        """reference SDXL/network.py:233-256"""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        super().reset_parameters()
        self._invalidate()

    def build_condition_input(self, resolution: int) -> torch.Tensor:
        """reference SDXL/network.py:325-332: original_size + crop_top_left + target_size"""
        return torch.tensor([[resolution, resolution, 0, 0, resolution, resolution]], dtype=torch.float32)

    def _invalidate(self):
        self._packed, self._plans, self._ctx = None, {}, None

    def _apply(self, fn, *a, **kw):
        self._invalidate()
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, state_dict: Mapping[str, Any], **kwargs):
        self._invalidate()
        return super().load_state_dict(state_dict, **kwargs)

    # ---- weight repacking --------------------------------------------------------------------------------
    def _resnets(self) -> List[Tuple[str, _Resnet]]:
        u = self.unet
        out = []
        for i, b in enumerate(u.down_blocks):
            out += [(f"down.{i}.{j}", r) for j, r in enumerate(b.resnets)]
        out += [(f"mid.{j}", r) for j, r in enumerate(u.mid_block.resnets)]
        for i, b in enumerate(u.up_blocks):
            out += [(f"up.{i}.{j}", r) for j, r in enumerate(b.resnets)]
        return out

    def _transformers(self) -> List[Tuple[str, _Transformer2D]]:
        u = self.unet
        out = []
        for i, b in enumerate(u.down_blocks):
            if hasattr(b, "attentions"):
                out += [(f"down.{i}.{j}", a) for j, a in enumerate(b.attentions)]
        out.append(("mid.0", u.mid_block.attentions[0]))
        for i, b in enumerate(u.up_blocks):
            if hasattr(b, "attentions"):
                out += [(f"up.{i}.{j}", a) for j, a in enumerate(b.attentions)]
        return out

    @torch.no_grad()
    def repack(self) -> None:
        u = self.unet
        dev = u.conv_out.weight.device
        if dev.type != "cuda" and not _ALLOW_HOST_PLANS:
            raise L.FastGenB200Error("StableDiffusionXL (B200) must live on a CUDA device: there is no CPU path")
        L.lib()
        bf, f32 = torch.bfloat16, torch.float32

        def conv_w(c: nn.Conv2d, cin_pad: Optional[int] = None) -> torch.Tensor:   # [cout][tap][cin]
            w = c.weight.detach().permute(0, 2, 3, 1)
            if cin_pad is not None and cin_pad != w.shape[-1]:
                w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[-1]))
            return w.reshape(w.shape[0], -1).to(bf).contiguous()

        w = lambda lin: lin.weight.detach().to(bf).contiguous()                     # noqa: E731
        b = lambda m: m.bias.detach().to(f32).contiguous()                          # noqa: E731
        pk: Dict[str, Any] = {}
        pk["conv_in.w"], pk["conv_in.b"] = conv_w(u.conv_in, 8), b(u.conv_in)
        for nm, mod in (("te", u.time_embedding),) + ((("ae", u.add_embedding),) if self._has_add else ()):
            pk[nm + "1.w"], pk[nm + "1.b"] = w(mod.linear_1), b(mod.linear_1)
            pk[nm + "2.w"], pk[nm + "2.b"] = w(mod.linear_2), b(mod.linear_2)
        # the time projections of every ResnetBlock2D as ONE GEMM on SiLU(emb)
        off, tw, tb = 0, [], []
        for key, r in self._resnets():
            pk[key + ".temb_off"] = off
            off += r.cout
            tw.append(r.time_emb_proj.weight)
            tb.append(r.time_emb_proj.bias)
            pk[key + ".n1"] = (r.norm1.weight.detach().to(f32).contiguous(), b(r.norm1))
            pk[key + ".n2"] = (r.norm2.weight.detach().to(f32).contiguous(), b(r.norm2))
            pk[key + ".c1.w"], pk[key + ".c1.b"] = conv_w(r.conv1), b(r.conv1)
            pk[key + ".c2.w"] = conv_w(r.conv2)
            if hasattr(r, "conv_shortcut"):
                pk[key + ".sc.w"] = r.conv_shortcut.weight.detach().reshape(r.cout, r.cin).to(bf).contiguous()
                pk[key + ".c2.b"] = (r.conv2.bias.detach().to(f32) + r.conv_shortcut.bias.detach().to(f32)).contiguous()
            else:
                pk[key + ".c2.b"] = b(r.conv2)
        pk["temb.w"] = torch.cat(tw, 0).detach().to(bf).contiguous()
        pk["temb.b"] = torch.cat(tb, 0).detach().to(f32).contiguous()
        for key, a in self._transformers():
            k = "t." + key
            pk[k + ".norm"] = (a.norm.weight.detach().to(f32).contiguous(), b(a.norm))
            w2 = lambda m: m.weight.detach().reshape(m.weight.shape[0], -1).to(bf).contiguous()   # noqa: E731  Linear | 1x1 conv
            pk[k + ".in.w"], pk[k + ".in.b"] = w2(a.proj_in), b(a.proj_in)
            pk[k + ".out.w"], pk[k + ".out.b"] = w2(a.proj_out), b(a.proj_out)
            for li, blk in enumerate(a.transformer_blocks):
                q = f"{k}.{li}"
                for n_ in ("norm1", "norm2", "norm3"):
                    ln = getattr(blk, n_)
                    pk[f"{q}.{n_}"] = (ln.weight.detach().to(f32).contiguous(), b(ln))
                pk[q + ".qkv.w"] = torch.cat([blk.attn1.to_q.weight, blk.attn1.to_k.weight, blk.attn1.to_v.weight],
                                             0).detach().to(bf).contiguous()
                pk[q + ".o1.w"], pk[q + ".o1.b"] = w(blk.attn1.to_out[0]), b(blk.attn1.to_out[0])
                pk[q + ".q2.w"] = w(blk.attn2.to_q)
                pk[q + ".kv2.w"] = torch.cat([blk.attn2.to_k.weight, blk.attn2.to_v.weight], 0).detach().to(bf).contiguous()
                pk[q + ".o2.w"], pk[q + ".o2.b"] = w(blk.attn2.to_out[0]), b(blk.attn2.to_out[0])
                # GEGLU fused into the ff.net.0 GEMM tail: rows interleaved to (hidden_j, gate_j) pairs (FG_ACT_GEGLU)
                fp = blk.ff.net[0].proj
                inner = fp.out_features // 2
                perm = torch.stack([torch.arange(inner), torch.arange(inner) + inner], 1).reshape(-1).to(fp.weight.device)
                pk[q + ".ff0.w"] = fp.weight.detach()[perm].to(bf).contiguous()
                pk[q + ".ff0.b"] = fp.bias.detach()[perm].to(f32).contiguous()
                pk[q + ".ff2.w"], pk[q + ".ff2.b"] = w(blk.ff.net[2]), b(blk.ff.net[2])
        for i, blk in enumerate(u.down_blocks):
            if hasattr(blk, "downsamplers"):
                pk[f"ds.{i}.w"], pk[f"ds.{i}.b"] = conv_w(blk.downsamplers[0].conv), b(blk.downsamplers[0].conv)
        for i, blk in enumerate(u.up_blocks):
            if hasattr(blk, "upsamplers"):
                pk[f"us.{i}.w"], pk[f"us.{i}.b"] = conv_w(blk.upsamplers[0].conv), b(blk.upsamplers[0].conv)
        pk["norm_out"] = (u.conv_norm_out.weight.detach().to(f32).contiguous(), b(u.conv_norm_out))
        pk["conv_out.w"], pk["conv_out.b"] = conv_w(u.conv_out), b(u.conv_out)
        self._packed, self._plans, self._ctx = pk, {}, None

    # ---- text keys / values of every cross-attention layer (static per prompt) ------------------------------
    def _context(self, text: torch.Tensor, mask: Optional[torch.Tensor] = None, masked: bool = False) -> Dict[str, Any]:
        """text keys / values of every cross-attention layer, computed once per condition.  ``masked`` (SD 1.5): the
        cross-attentions take per-sample key counts from ``ent["lens"]``, filled from the prefix ``encoder_attention_mask``
        ``mask`` [B, L] of 0 / 1 (None: all keys)."""
        key = tensor_identity(text)
        if key is not None and mask is not None:
            mkey = tensor_identity(mask)
            key = None if mkey is None else key + mkey
        B, Lt, X = text.shape
        if X != self.cfg["cross_attention_dim"]:
            raise ValueError(f"text embedding width {X} != cross_attention_dim {self.cfg['cross_attention_dim']}")
        ent = self._ctx
        if ent is None or ent["batch"] != B or ent["len"] != Lt:
            dev, bf, pk = text.device, torch.bfloat16, self._packed
            ent = {"key": None, "batch": B, "len": Lt, "text": torch.empty(B * Lt, X, device=dev, dtype=bf), "kv": {}}
            if masked:
                ent["lens"] = torch.full((B,), Lt, device=dev, dtype=torch.int32)
            plan = _Plan()
            for tk, a in self._transformers():
                for li in range(len(a.transformer_blocks)):
                    q = f"t.{tk}.{li}"
                    kv = torch.empty(B * Lt, 2 * a.dim, device=dev, dtype=bf)
                    ent["kv"][q] = kv
                    p = ops.conv_gemm_params([ops.Seg(ent["text"], X, pk[q + ".kv2.w"], X, 1)], B * Lt, 1, 1, 2 * a.dim,
                                             kv, 2 * a.dim)
                    fn, args, pkeep = ops.conv_gemm_launch(p)
                    plan.keep.append((p, pkeep))
                    plan.add(fn, *args, desc="kv2", flops=2.0 * B * Lt * X * 2 * a.dim)
            ent["plan"] = plan
            self._ctx = ent
            self._plans = {}     # forward plans point into the previous buffers
        if key is None or ent["key"] != key:
            ent["text"].copy_(text.reshape(B * Lt, X))
            ent["plan"].run(torch.cuda.current_stream().cuda_stream, self.profile)
            if masked:
                if mask is None:
                    ent["lens"].fill_(Lt)
                else:
                    if tuple(mask.shape) != (B, Lt):
                        raise ValueError(f"encoder_attention_mask shape {tuple(mask.shape)} != {(B, Lt)}")
                    m = mask.to(device=text.device) != 0
                    lens = m.sum(dim=1)
                    # the kernel takes key COUNTS: the mask must keep a prefix (what a tokenizer's padding mask is) and at
                    # least one key per sample -- one host check per new condition
                    prefix = bool((m[:, :-1] | ~m[:, 1:]).all().item()) if Lt > 1 else True
                    if not prefix or int(lens.min().item()) < 1:
                        raise NotImplementedError("encoder_attention_mask must keep a non-empty prefix of the text keys "
                                                  "(a padding mask); general masks are not supported")
                    ent["lens"].copy_(lens.to(torch.int32))
            ent["key"], ent["keep"] = key, (text, mask)
        return ent

    # ---- the launch list -------------------------------------------------------------------------------------
    def _plan(self, B: int, H: int, W: int, x_dtype: torch.dtype, ctx: Dict[str, Any]):
        pkey = (B, H, W, x_dtype)
        hit = self._plans.get(pkey)
        if hit is not None:
            return hit
        cfg, pk, lib, u = self.cfg, self._packed, L.lib(), self.unet
        lay = layout(cfg)
        dev = u.conv_out.weight.device
        bf, f32 = torch.bfloat16, torch.float32
        G, T, Lt = cfg["norm_groups"], lay["temb"], ctx["len"]
        nlev = len(cfg["block_out_channels"])
        if H % (1 << (nlev - 1)) or W % (1 << (nlev - 1)):
            raise ValueError(f"latent size {H}x{W} must be divisible by {1 << (nlev - 1)}")
        plan = _Plan()
        keep = plan.keep
        new = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)   # noqa: E731

        # ---- scratch: one flat buffer per role, sized for the largest user (layers run in order on one stream)
        need: Dict[Tuple[str, torch.dtype], int] = {}

        def req(name: str, numel: int, dt=bf):
            need[(name, dt)] = max(need.get((name, dt), 0), numel)

        def sizes():
            h, w = H, W
            for i, d in enumerate(lay["down"]):
                for ci, co in d["res"]:
                    yield ("res", h, w, ci, co)
                    if d["attn"]:
                        yield ("tr", h, w, co, co)
                if d["down"]:
                    yield ("ds", h, w, d["c"], d["c"])
                    h, w = h // 2, w // 2
            yield ("res", h, w, lay["mid_c"], lay["mid_c"])
            yield ("tr", h, w, lay["mid_c"], lay["mid_c"])
            for i, uu in enumerate(lay["up"]):
                for rin, skip, co in uu["res"]:
                    yield ("res", h, w, rin + skip, co)
                    if uu["attn"]:
                        yield ("tr", h, w, co, co)
                if uu["up"]:
                    yield ("us", h, w, uu["c"], uu["c"])
                    h, w = 2 * h, 2 * w
            yield ("res", h, w, cfg["block_out_channels"][0], 8)     # output head reuses h0

        for kind, h, w, ci, co in sizes():
            px = B * h * w
            if kind == "res":
                req("h0", px * ci), req("raw", px * ci), req("h1", px * co, f32), req("h2", px * co), req("mid", px * co, f32)
            elif kind == "tr":
                req("nrm", px * ci), req("hs", px * ci, f32), req("qkv", px * 3 * ci), req("att", px * ci)
                req("gg", px * 4 * ci)
            elif kind == "ds":
                req("h0", px * ci), req("full", px * co, f32)
            else:
                req("h0", 4 * px * ci)
        pool = {k: new(v, dt=k[1]) for k, v in need.items()}
        keep.append(pool)

        def scratch(name: str, shape: Sequence[int], dt=bf) -> torch.Tensor:
            n_ = 1
            for s in shape:
                n_ *= s
            return pool[(name, dt)][:n_].view(*shape)

        n_norms = 2 * len(self._resnets()) + len(self._transformers()) + 2
        arena = new(n_norms, B, G, 2, dt=torch.float64)
        used = [0]
        plan.add(lib.fg_zero, arena.data_ptr(), arena.numel() * 8, desc="zero_stats")

        def new_stats() -> torch.Tensor:
            i = used[0]
            used[0] += 1
            assert i < arena.shape[0], "planner bug: statistics arena too small"
            return arena[i]

        def can_fuse(px: int, cout: int, out_ld: int) -> bool:
            return px >= 32 and cout % 4 == 0 and cout > 16 and out_ld % 4 == 0

        def fused_stats(h: int, w: int, cout: int, out_ld: int, c_norm: Optional[int]):
            """conv kwargs that accumulate the statistics of a norm over exactly this tensor, or None"""
            if c_norm is None or cout != c_norm or not can_fuse(h * w, cout, out_ld):
                return None
            return dict(stats=new_stats(), stats_groups=G, stats_cg=c_norm // G, stats_ch0=0)

        def stats_kw(dst: "_Act", own_c: Optional[int], cat: Optional[Tuple[torch.Tensor, int, int]]):
            """statistics targets of a layer output: its own next norm (over own_c == dst.c channels) and / or the norm
            of the skip concatenation it is part of -- cat = (slot, channels per group, channel offset).  Returns
            (conv kwargs, slot of the own norm or None)."""
            kw: Dict[str, Any] = {}
            own = fused_stats(dst.h, dst.w, dst.c, dst.ld, own_c)
            if own:
                kw.update(own)
            if cat is not None:
                slot, cg, ch0 = cat
                if own:
                    kw.update(stats2=slot, stats2_groups=G, stats2_cg=cg, stats2_ch0=ch0)
                else:
                    kw.update(stats=slot, stats_groups=G, stats_cg=cg, stats_ch0=ch0)
            return kw, (own["stats"] if own else None)

        def conv(segs, n, h, w, cout, out, out_ld, desc, **kw):
            p = ops.conv_gemm_params(segs, n, h, w, cout, out, out_ld, **kw)
            fn, args, pkeep = ops.conv_gemm_launch(p)
            keep.append((p, pkeep, segs, out, kw))
            plan.add(fn, *args, desc=desc, flops=2.0 * n * h * w * cout * sum(s.cin * s.taps for s in segs))

        def gemm(a, k, wt, rows, cout, out, out_ld, desc, **kw):
            conv([ops.Seg(a, a.stride(0), wt, k, 1)], rows, 1, 1, cout, out, out_ld, desc, **kw)

        def gn(x: _Act, gb, eps: float, y: torch.Tensor, silu: bool, stats: Optional[torch.Tensor], raw=None):
            xv = x.t
            if stats is None:
                stats = new_stats()
                plan.add(lib.fg_gn_stats, xv.data_ptr(), ops._DT[xv.dtype], x.ld, B, x.h * x.w, x.c, G, stats.data_ptr(),
                         desc="gn_stats")
            p = ops.gn_apply_params(xv, x.ld, B, x.h, x.w, x.c, G, stats, eps, gb[0], gb[1], y, x.c, silu=silu, raw=raw,
                                    raw_ld=x.c)
            keep.append((p, xv, y, raw, stats))
            plan.add(lib.fg_gn_apply, C.byref(p), desc="gn_apply")

        def ln(x, rows, d, gb, y, desc):
            p = ops.ln_mod_params(x, d, rows, d, 1e-5, y, d, gamma=gb[0], beta=gb[1])
            keep.append((p, x, y))
            plan.add(lib.fg_ln_modulate, C.byref(p), desc=desc)

        def attn(q, q_ld, k, v, kv_ld, kv_len, tq, heads, d, out, out_ld, desc, lens=None):
            p = ops.flash_attn_params(q, q_ld, tq * q_ld, [(k, v, kv_ld, kv_len * kv_ld, kv_len)], B, heads, d, tq, out,
                                      out_ld, tq * out_ld, kv_len0_per_sample=lens)
            wsz = ops.flash_attn_workspace_size(p)
            if wsz > 0:
                ws_ = pool.get(("attn_ws", torch.uint8))
                if ws_ is None or ws_.numel() < wsz:
                    ws_ = torch.empty(wsz, device=dev, dtype=torch.uint8)
                    pool[("attn_ws", torch.uint8)] = ws_
                p.workspace, p.workspace_bytes = ws_.data_ptr(), ws_.numel()
                keep.append(ws_)
            keep.append((p, q, k, v, out))
            plan.add(lib.fg_flash_attn, C.byref(p), desc=desc, flops=4.0 * B * tq * kv_len * heads * d)

        # ---- conditioning path ------------------------------------------------------------------------------------
        c0 = cfg["block_out_channels"][0]
        ws: Dict[str, torch.Tensor] = dict(
            x_in=new(B, cfg["in_channels"], H, W, dt=x_dtype), xin=new(B, H, W, 8, dt=bf), t_sin=new(B, c0, dt=bf),
            add_in=new(B, max(lay["proj_in"], 8), dt=bf), e1=new(B, T, dt=bf), emb_t=new(B, T), a1=new(B, T, dt=bf), emb=new(B, T),
            emb_act=new(B, T, dt=bf), temb=new(B, pk["temb.w"].shape[0]), f_out=new(B, H, W, 4),
            eps=new(B, cfg["out_channels"], H, W, dt=x_dtype))
        keep.append(ws)
        plan.add(lib.fg_edm_input, ws["x_in"].data_ptr(), ops._DT[x_dtype], B, cfg["in_channels"], H * W, None,
                 ws["xin"].data_ptr(), None, 8, desc="nchw_to_nhwc")
        gemm(ws["t_sin"], c0, pk["te1.w"], B, T, ws["e1"], T, "time1", bias=pk["te1.b"], act=L.FG_ACT_SILU)
        if self._has_add:
            gemm(ws["e1"], T, pk["te2.w"], B, T, ws["emb_t"], T, "time2", bias=pk["te2.b"])
            gemm(ws["add_in"], lay["proj_in"], pk["ae1.w"], B, T, ws["a1"], T, "add1", bias=pk["ae1.b"], act=L.FG_ACT_SILU)
            gemm(ws["a1"], T, pk["ae2.w"], B, T, ws["emb"], T, "add2", bias=pk["ae2.b"], res=ws["emb_t"], res_ld=T,
                 res_mode=L.FG_RES_SAME)
        else:   # SD 1.5: the time embedding alone
            gemm(ws["e1"], T, pk["te2.w"], B, T, ws["emb"], T, "time2", bias=pk["te2.b"])
        plan.add(lib.fg_split_bf16, ws["emb"].data_ptr(), T, None, 0, B, T, 1, ws["emb_act"].data_ptr(), None, T, None, 0,
                 desc="silu_emb")
        tl = ws["temb"].shape[1]
        gemm(ws["emb_act"], T, pk["temb.w"], B, tl, ws["temb"], tl, "time_proj_all", bias=pk["temb.b"])

        # ---- layers -----------------------------------------------------------------------------------------------
        def resnet(key: str, r: _Resnet, x: _Act, dst: _Act, x_stats, out_kw: Dict[str, Any]):
            """ResnetBlock2D; out_kw: statistics kwargs of the output conv (see stats_kw)"""
            h, w, cin, cout = x.h, x.w, r.cin, r.cout
            assert x.c == cin and dst.c == cout
            short = (key + ".sc.w") in pk
            h0 = scratch("h0", [B, h, w, cin])
            raw = scratch("raw", [B, h, w, cin]) if short else None
            gn(x, pk[key + ".n1"], 1e-5, h0, True, x_stats, raw=raw)
            h1 = scratch("h1", [B, h, w, cout], f32)
            st1 = fused_stats(h, w, cout, cout, cout)
            conv([ops.Seg(h0, cin, pk[key + ".c1.w"], cin, 9)], B, h, w, cout, h1, cout, "res.conv1", bias=pk[key + ".c1.b"],
                 rowvec=ws["temb"][:, pk[key + ".temb_off"]:], rowvec_ld=tl, **(st1 or {}))
            h2 = scratch("h2", [B, h, w, cout])
            gn(_Act(h1, 0, cout), pk[key + ".n2"], 1e-5, h2, True, st1["stats"] if st1 else None)
            segs = [ops.Seg(h2, cout, pk[key + ".c2.w"], cout, 9)]
            kw: Dict[str, Any] = {}
            if short:
                segs.append(ops.Seg(raw, cin, pk[key + ".sc.w"], cin, 1))
            else:
                kw = dict(res=x.t, res_ld=x.ld, res_mode=L.FG_RES_SAME)
            conv(segs, B, h, w, cout, dst.t, dst.ld, "res.conv2", bias=pk[key + ".c2.b"], **kw, **out_kw)

        def transformer(key: str, a: _Transformer2D, x: _Act, dst: _Act, x_stats, out_kw: Dict[str, Any]):
            h, w, c, heads = x.h, x.w, a.dim, a.heads
            d, N = c // heads, x.h * x.w
            rows = B * N
            k = "t." + key
            nrm, hs = scratch("nrm", [rows, c]), scratch("hs", [rows, c], f32)     # fp32 token (residual) stream
            qkv, att = scratch("qkv", [rows, 3 * c]), scratch("att", [rows, c])
            gg = scratch("gg", [rows, 4 * c])
            gn(x, pk[k + ".norm"], 1e-6, nrm.view(B, h, w, c), False, x_stats)
            gemm(nrm, c, pk[k + ".in.w"], rows, c, hs, c, "proj_in", bias=pk[k + ".in.b"])
            for li in range(len(a.transformer_blocks)):
                q = f"{k}.{li}"
                ln(hs, rows, c, pk[q + ".norm1"], nrm, "ln1")
                gemm(nrm, c, pk[q + ".qkv.w"], rows, 3 * c, qkv, 3 * c, "qkv")
                attn(qkv, 3 * c, qkv[:, c:], qkv[:, 2 * c:], 3 * c, N, N, heads, d, att, c, "self_attn")
                gemm(att, c, pk[q + ".o1.w"], rows, c, hs, c, "attn1_out", bias=pk[q + ".o1.b"], res=hs, res_ld=c,
                     res_mode=L.FG_RES_SAME)
                ln(hs, rows, c, pk[q + ".norm2"], nrm, "ln2")
                gemm(nrm, c, pk[q + ".q2.w"], rows, c, qkv, 3 * c, "q2")
                kv = ctx["kv"][q]
                attn(qkv, 3 * c, kv, kv[:, c:], 2 * c, Lt, N, heads, d, att, c, "cross_attn", lens=ctx.get("lens"))
                gemm(att, c, pk[q + ".o2.w"], rows, c, hs, c, "attn2_out", bias=pk[q + ".o2.b"], res=hs, res_ld=c,
                     res_mode=L.FG_RES_SAME)
                ln(hs, rows, c, pk[q + ".norm3"], nrm, "ln3")
                gemm(nrm, c, pk[q + ".ff0.w"], rows, 8 * c, gg, 4 * c, "ff0_geglu", bias=pk[q + ".ff0.b"],
                     act=L.FG_ACT_GEGLU)
                gemm(gg, 4 * c, pk[q + ".ff2.w"], rows, c, hs, c, "ff2", bias=pk[q + ".ff2.b"], res=hs, res_ld=c,
                     res_mode=L.FG_RES_SAME)
            plan.add(lib.fg_cast_resample, hs.data_ptr(), L.FG_F32, c, rows, 1, 1, c, L.FG_RES_NONE, nrm.data_ptr(), c,
                     desc="cast_tokens")
            conv([ops.Seg(nrm.view(B, h, w, c), c, pk[k + ".out.w"], c, 1)], B, h, w, c, dst.t, dst.ld, "proj_out",
                 bias=pk[k + ".out.b"], res=x.t, res_ld=x.ld, res_mode=L.FG_RES_SAME, **out_kw)

        def cast(x: _Act, up: bool) -> torch.Tensor:
            s = 2 if up else 1
            y = scratch("h0", [B, s * x.h, s * x.w, x.c])
            xv = x.t
            plan.add(lib.fg_cast_resample, xv.data_ptr(), ops._DT[xv.dtype], x.ld, B, x.h, x.w, x.c,
                     L.FG_RES_UP2 if up else L.FG_RES_NONE, y.data_ptr(), x.c, desc="cast_up2" if up else "cast")
            keep.append((xv, y))
            return y

        # ---- activation homes: every skip producer writes into the concat buffer of its consumer -------------------
        # walk the down path once to learn (resolution, channels) of the 1 + sum(...) skips, then allocate the
        # concat buffers of the up path in pop order
        skip_shapes: List[Tuple[int, int, int]] = [(H, W, c0)]
        h, w = H, W
        for d in lay["down"]:
            skip_shapes += [(h, w, d["c"])] * len(d["res"])
            if d["down"]:
                h, w = h // 2, w // 2
                skip_shapes.append((h, w, d["c"]))
        pending = list(range(len(skip_shapes)))
        skip_home: Dict[int, _Act] = {}
        skip_user: Dict[int, Tuple[int, int]] = {}       # skip index -> (up block, resnet) that concatenates it
        cat_of: Dict[Tuple[int, int], torch.Tensor] = {}
        for i, uu in enumerate(lay["up"]):
            for j, (rin, skip, co) in enumerate(uu["res"]):
                s = pending.pop()
                sh, sw, sc = skip_shapes[s]
                assert sc == skip, "planner bug: skip channel mismatch"
                buf = new(B, sh, sw, rin + skip)
                cat_of[(i, j)] = buf
                skip_home[s] = _Act(buf, rin, skip)
                skip_user[s] = (i, j)
        assert not pending
        # GroupNorm statistics of a concatenation come from its two producers (fg_conv_gemm stats / stats2 with a channel
        # offset) when both can fuse them; a skip written by a downsampler (stride-1 conv + subsampling) cannot
        down_skips = set()
        k_ = 0
        for d in lay["down"]:
            k_ += len(d["res"])
            if d["down"]:
                k_ += 1
                down_skips.add(k_)
        cat_slot: Dict[Tuple[int, int], torch.Tensor] = {}
        for s_, (i, j) in skip_user.items():
            buf = cat_of[(i, j)]
            rin, skip, _ = lay["up"][i]["res"][j]
            px, ld = buf.shape[1] * buf.shape[2], buf.shape[3]
            if s_ not in down_skips and can_fuse(px, rin, ld) and can_fuse(px, skip, ld) and (rin + skip) % G == 0:
                cat_slot[(i, j)] = new_stats()

        def cat_target(key: Optional[Tuple[int, int]], part: str):
            """(slot, channels per group, channel offset) of concat norm `key` for its x part / skip part, or None"""
            if key is None or key not in cat_slot:
                return None
            rin, skip, _ = lay["up"][key[0]]["res"][key[1]]
            return cat_slot[key], (rin + skip) // G, (0 if part == "x" else rin)

        # ---- down path --------------------------------------------------------------------------------------------
        si = 0
        x = skip_home[si]
        kw0, x_stats = stats_kw(x, c0, cat_target(skip_user[si], "skip"))
        conv([ops.Seg(ws["xin"], 8, pk["conv_in.w"], 8, 9)], B, H, W, c0, x.t, x.ld, "conv_in", bias=pk["conv_in.b"], **kw0)
        si += 1
        for i, d in enumerate(lay["down"]):
            blk = u.down_blocks[i]
            for j in range(len(d["res"])):
                dst = skip_home[si]
                last = j == len(d["res"]) - 1
                after_c = (None if d["down"] else lay["mid_c"]) if last else d["c"]   # norm that reads dst next
                okw, own = stats_kw(dst, after_c, cat_target(skip_user[si], "skip"))
                si += 1
                if d["attn"]:
                    mid = _Act(scratch("mid", [B, x.h, x.w, d["c"]], f32), 0, d["c"])
                    mkw, s_mid = stats_kw(mid, d["c"], None)
                    resnet(f"down.{i}.{j}", blk.resnets[j], x, mid, x_stats, mkw)
                    transformer(f"down.{i}.{j}", blk.attentions[j], mid, dst, s_mid, okw)
                else:
                    resnet(f"down.{i}.{j}", blk.resnets[j], x, dst, x_stats, okw)
                x, x_stats = dst, own
            if d["down"]:
                dst = skip_home[si]
                si += 1
                xb = cast(x, False)
                full = scratch("full", [B, x.h, x.w, d["c"]], f32)
                conv([ops.Seg(xb, d["c"], pk[f"ds.{i}.w"], d["c"], 9)], B, x.h, x.w, d["c"], full, d["c"], "downsample",
                     bias=pk[f"ds.{i}.b"])
                ho, wo, c = x.h // 2, x.w // 2, d["c"]
                args, kp = ops.strided_copy_args(full, dst.t, (B, ho, wo, 1),
                                                 (x.h * x.w * c, 2 * x.w * c, 2 * c, 0),
                                                 (ho * wo * dst.ld, wo * dst.ld, dst.ld, 0), c)
                keep.append(kp)
                plan.add(lib.fg_strided_copy, *args, desc="keep_even")
                x, x_stats = dst, None
        # ---- mid block --------------------------------------------------------------------------------------------
        mc = lay["mid_c"]
        m1 = _Act(new(B, x.h, x.w, mc), 0, mc)
        kw1, s1 = stats_kw(m1, mc, None)
        resnet("mid.0", u.mid_block.resnets[0], x, m1, x_stats, kw1)
        m2 = _Act(new(B, x.h, x.w, mc), 0, mc)
        kw2, s2 = stats_kw(m2, mc, None)
        transformer("mid.0", u.mid_block.attentions[0], m1, m2, s1, kw2)
        x = _Act(cat_of[(0, 0)], 0, mc)
        kw3, _ = stats_kw(x, None, cat_target((0, 0), "x"))
        resnet("mid.1", u.mid_block.resnets[1], m2, x, s2, kw3)
        # ---- up path ----------------------------------------------------------------------------------------------
        nb = len(lay["up"])
        for i, uu in enumerate(lay["up"]):
            blk = u.up_blocks[i]
            nres = len(uu["res"])
            for j, (rin, skip, co) in enumerate(uu["res"]):
                cat = cat_of[(i, j)]
                assert x.buf is cat and x.off == 0 and x.c == rin, "planner bug: up-path input is not in its concat buffer"
                xin = _Act(cat, 0, rin + skip)
                xin_stats = cat_slot.get((i, j))      # None: this norm runs its own fg_gn_stats pass
                last = j == nres - 1
                if not last:
                    dst = _Act(cat_of[(i, j + 1)], 0, co)
                    okw, own = stats_kw(dst, None, cat_target((i, j + 1), "x"))
                elif uu["up"]:
                    dst = _Act(new(B, xin.h, xin.w, co), 0, co)
                    okw, own = {}, None
                else:
                    dst = _Act(new(B, xin.h, xin.w, co), 0, co)
                    okw, own = stats_kw(dst, co, None)                 # conv_norm_out reads it
                if uu["attn"]:
                    mid = _Act(scratch("mid", [B, xin.h, xin.w, co], f32), 0, co)
                    mkw, s_mid = stats_kw(mid, co, None)
                    resnet(f"up.{i}.{j}", blk.resnets[j], xin, mid, xin_stats, mkw)
                    transformer(f"up.{i}.{j}", blk.attentions[j], mid, dst, s_mid, okw)
                else:
                    resnet(f"up.{i}.{j}", blk.resnets[j], xin, dst, xin_stats, okw)
                x, x_stats = dst, own
            if uu["up"]:
                xu = cast(x, True)
                dst = _Act(cat_of[(i + 1, 0)], 0, uu["c"])
                ukw, _ = stats_kw(dst, None, cat_target((i + 1, 0), "x"))
                conv([ops.Seg(xu, uu["c"], pk[f"us.{i}.w"], uu["c"], 9)], B, 2 * x.h, 2 * x.w, uu["c"], dst.t, dst.ld,
                     "upsample", bias=pk[f"us.{i}.b"], **ukw)
                x, x_stats = dst, None
        assert i == nb - 1
        # ---- output head --------------------------------------------------------------------------------------------
        hb = scratch("h0", [B, H, W, c0])
        gn(x, pk["norm_out"], 1e-5, hb, True, x_stats)
        conv([ops.Seg(hb, c0, pk["conv_out.w"], c0, 9)], B, H, W, cfg["out_channels"], ws["f_out"], 4, "conv_out",
             bias=pk["conv_out.b"])
        stp = L.EdmStepParams()
        stp.f, stp.f_ld = ws["f_out"].data_ptr(), 4
        stp.x_t, stp.x_dtype = ws["x_in"].data_ptr(), ops._DT[x_dtype]
        stp.n, stp.c, stp.hw = B, cfg["out_channels"], H * W
        stp.x0 = ws["eps"].data_ptr()
        keep.append(stp)
        plan.add(lib.fg_edm_step, C.byref(stp), desc="nhwc_to_nchw")
        self._plans[pkey] = (plan, ws)
        return plan, ws

    # ---- forward (reference SDXL/network.py:334-416) -------------------------------------------------------------
    def forward(self, x_t: torch.Tensor, t: torch.Tensor, condition: Optional[List[torch.Tensor]] = None,
                r: Optional[torch.Tensor] = None, return_features_early: bool = False,
                feature_indices: Optional[Set[int]] = None, return_logvar: bool = False,
                fwd_pred_type: Optional[str] = None, **fwd_kwargs):
        if r is not None:
            raise NotImplementedError(f"r is not yet supported for {'SDXL' if self._has_add else 'SD15'}")
        if feature_indices is None:
            feature_indices = {}
        if return_features_early and len(feature_indices) == 0:
            return []
        if fwd_pred_type is None:
            fwd_pred_type = self.net_pred_type
        else:
            assert fwd_pred_type in NET_PRED_TYPES, f"{fwd_pred_type} is not supported as fwd_pred_type"
        if return_features_early or len(feature_indices) or return_logvar:
            raise NotImplementedError("feature / logvar outputs are training-time paths of the reference")
        if not x_t.is_cuda:
            raise L.FastGenB200Error("StableDiffusionXL (B200) needs CUDA tensors: there is no CPU path")
        if x_t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"x_t must be float32 or bfloat16, got {x_t.dtype}")
        if condition is None or len(condition) != 2:
            raise ValueError("condition must be [prompt_embeds [B,L,X], pooled_prompt_embeds [B,P]]")
        with torch.no_grad():
            if self._packed is None:
                self.repack()
            text, pooled = condition[0], condition[1]
            B, Cc, H, W = x_t.shape
            if Cc != self.cfg["in_channels"]:
                raise ValueError(f"x_t has {Cc} channels, the UNet takes {self.cfg['in_channels']}")
            if text.shape[0] != B or (pooled is not None and pooled.shape[0] != B):
                raise ValueError("condition batch size does not match x_t")
            if self._has_add:
                ctx = self._context(text)
            else:   # SD 1.5: condition = [encoder_hidden_states, encoder_attention_mask]
                ctx = self._context(text, mask=pooled, masked=True)
            plan, ws = self._plan(B, H, W, x_t.dtype, ctx)
            dev = x_t.device
            timestep = self.noise_scheduler.rescale_t(torch.atleast_1d(t).to(dev))
            if timestep.numel() == 1 and B > 1:
                timestep = timestep.expand(B)
            ws["x_in"].copy_(x_t)
            ws["t_sin"].copy_(timestep_sinusoid(timestep, self.cfg["block_out_channels"][0]))
            if self._has_add:
                ids = self.add_time_ids.to(dev).repeat(B, 1)
                te = timestep_sinusoid(ids.flatten(), self.cfg["addition_time_embed_dim"]).reshape(B, -1)
                ws["add_in"].copy_(torch.cat([pooled.to(te.dtype), te], dim=-1))
            plan.run(torch.cuda.current_stream().cuda_stream, self.profile)
            out = ws["eps"].clone()
            return self.noise_scheduler.convert_model_output(x_t, out, torch.atleast_1d(t).to(dev),
                                                             src_pred_type=self.net_pred_type,
                                                             target_pred_type=fwd_pred_type)

    def sample(self, noise: torch.Tensor, condition: Optional[List[torch.Tensor]] = None,
               neg_condition: Optional[List[torch.Tensor]] = None, guidance_scale: Optional[float] = 7.5,
               num_steps: int = 50, **kwargs) -> torch.Tensor:
        """Teacher sampling: DDIM-style steps with optional classifier-free guidance -- reference
        ``StableDiffusionXL.sample`` (SDXL/network.py:418-473): same t list, same batched [negative | positive] eps
        prediction, ``x0 = eps_to_x0(x, eps, t)``, ``x <- alpha(t_next) x0 + sigma(t_next) eps``.  The elementwise passes
        are ``fg_axpby`` / ``fg_sub_div`` launches (fp64 coefficients, one rounding each)."""
        with torch.no_grad():
            sched = self.noise_scheduler
            t_list = sched.get_t_list(num_steps, device=noise.device)
            x = sched.latents(noise=noise, t_init=t_list[0])
            B = x.shape[0]
            cfg = guidance_scale is not None and guidance_scale > 1.0 and neg_condition is not None
            if cfg:
                cond_input = [torch.cat([neg_condition[0], condition[0]], dim=0),
                              torch.cat([neg_condition[1], condition[1]], dim=0)]
                g = torch.full((B,), float(guidance_scale), dtype=torch.float64, device=x.device)
            for t, t_next in zip(t_list[:-1], t_list[1:]):
                t_batch = t.expand(B)
                if cfg:
                    eps_pred = self(torch.cat([x, x], dim=0), torch.cat([t_batch, t_batch], dim=0), condition=cond_input,
                                    fwd_pred_type="eps")
                    eps_uncond, eps_cond = eps_pred.chunk(2)
                    eps_pred = ops.axpby(eps_uncond.contiguous(), 1.0 - g, eps_cond.contiguous(), g)
                else:
                    eps_pred = self(x, t_batch, condition=condition, fwd_pred_type="eps")
                x0_pred = sched.eps_to_x0(x, eps_pred, t_batch)
                x = sched.forward_process(x0_pred, eps_pred, t_next.expand(B))
            return x


class StableDiffusion15(StableDiffusionXL):
    """B200 drop-in for ``fastgen.networks.SD15.network.StableDiffusion15`` (SD15/network.py:438-680): the same
    ``UNet2DConditionModel`` driver (``classify_forward``, :46-341) at the SD 1.5 configuration -- four levels, 8 heads of
    40 / 80 / 160 / 160, 1x1-convolution proj_in / proj_out, no additional embedding -- with
    ``condition = [encoder_hidden_states [B, 77, 768], encoder_attention_mask [B, 77] | None]``.  The reference turns the
    mask into an additive ``-10000`` bias on the cross-attention logits (:142-145), whose softmax weights underflow to
    exactly 0; here a prefix mask (a tokenizer's padding mask) is passed to ``fg_flash_attn`` as per-sample key counts."""

    MODEL_ID_15 = "runwayml/stable-diffusion-v1-5"
    MODEL_ID_21 = "stabilityai/stable-diffusion-2-1"

    def __init__(self, model_id: str = MODEL_ID_15, net_pred_type: str = "eps", schedule_type: str = "sd",
                 load_pretrained: bool = True, unet_config: Optional[Mapping[str, Any]] = None, **model_kwargs):
        super().__init__(model_id=model_id, net_pred_type=net_pred_type, schedule_type=schedule_type,
                         load_pretrained=load_pretrained, unet_config=SD15_BASE if unet_config is None else unet_config,
                         **model_kwargs)
