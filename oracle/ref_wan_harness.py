"""Runs the UNMODIFIED reference Wan glue code on CPU against stand-in diffusers modules.

TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).  The reference's Wan networks are the diffusers
`WanTransformer3DModel` with its forward methods replaced by FastGen's own functions (fastgen/networks/Wan/network.py:836-851,
network_causal.py).  diffusers is absent here, so the third-party module internals (attention, RMSNorm, FP32LayerNorm,
rotary tables, embedders, feed-forward) are stand-ins built from oracle/wan_oracle.py's restatement -- but everything the
reference itself wrote runs as is: `classify_forward`, `classify_forward_prepare`, `classify_forward_block_forward`,
`block_forward` (AdaLN modulation, gates, r embedder, norm_temb, encoder_depth, head, unpatchify) and, for the causal
network, `CausalWanAttnProcessor.__call__`, `_wan_block_forward_inline_cache`, the external KV caches, the rotary time
offset and the block-causal mask.  oracle/make_golden.py:wan_glue_fixture() stores its outputs; tests/test_oracle_golden.py
checks oracle/wan_oracle.py against them, which pins the oracle's restatement of the REFERENCE'S code (the diffusers
internals stay "restated from the published sources").
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from oracle import wan_oracle as WO
from oracle.ref_shims import REF_ROOT, install_shims


# ---------------------------------------------------------------------------------------------------------
# stand-ins for the diffusers modules (math = oracle/wan_oracle.py)
# ---------------------------------------------------------------------------------------------------------
class FP32LayerNorm(nn.Module):
    def __init__(self, dim, eps, affine):
        super().__init__()
        self.eps = eps
        if affine:
            self.weight, self.bias = nn.Parameter(torch.ones(dim)), nn.Parameter(torch.zeros(dim))
        else:
            self.weight = self.bias = None

    def forward(self, x):
        return WO.fp32_layer_norm(x, self.weight, self.bias, self.eps)


class RMSNormAcrossHeads(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.eps, self.weight = eps, nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return WO.rms_norm(x, self.weight, self.eps)


class WanRotaryPosEmbed(nn.Module):
    """freqs_cos / freqs_sin tables [max_seq_len, head_dim]; forward gives the (cos, sin) of a [B,C,F,H,W] grid"""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        fc, fs = WO.rope_tables(cfg)
        self.register_buffer("freqs_cos", fc, persistent=False)
        self.register_buffer("freqs_sin", fs, persistent=False)
        self.patch_size = cfg["patch"]
        self.attention_head_dim = cfg["head_dim"]

    def forward(self, hidden_states):
        _, _, f, h, w = hidden_states.shape
        pt, ph, pw = self.cfg["patch"]
        return WO.rope_for_grid(self.cfg, f // pt, h // ph, w // pw, 0)

    def _apply(self, fn, *a, **kw):
        """keep the tables in float32 under `.to(bfloat16)`.  OPEN QUESTION (cannot be settled without diffusers 0.35.1):
        torch's Module.to casts floating-point buffers, so if diffusers does not special-case them a bf16 reference model
        rotates with bf16-rounded cos / sin tables (measured here: 4e-3 relative difference on the block output); the
        oracle and the CUDA path use float32 tables."""
        fc, fs = self.freqs_cos, self.freqs_sin
        super()._apply(fn, *a, **kw)
        dev = self.freqs_cos.device
        self.freqs_cos, self.freqs_sin = fc.to(dev), fs.to(dev)
        return self


class WanAttention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg["dim"]
        self.heads = cfg["heads"]
        self.to_q, self.to_k, self.to_v = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.to_out = nn.ModuleList([nn.Linear(d, d), nn.Dropout(0.0)])
        self.norm_q, self.norm_k = RMSNormAcrossHeads(d, cfg["eps"]), RMSNormAcrossHeads(d, cfg["eps"])
        self.add_k_proj = self.add_v_proj = self.norm_added_k = None
        self.fused_projections = False
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, rotary_emb=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states, attention_mask, rotary_emb, **kwargs)


class WanAttnProcessor:
    """diffusers' bidirectional processor restated: q/k/v projections, RMSNorm across heads, rotary, SDPA, to_out"""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, rotary_emb=None):
        q, k, v = _get_qkv_projections(attn, hidden_states, encoder_hidden_states)
        q, k = attn.norm_q(q), attn.norm_k(k)
        q, k, v = (t.unflatten(2, (attn.heads, -1)) for t in (q, k, v))
        if rotary_emb is not None:
            q, k = WO.apply_rotary_emb(q, *rotary_emb), WO.apply_rotary_emb(k, *rotary_emb)
        o = dispatch_attention_fn(q, k, v, attn_mask=attention_mask)
        o = o.flatten(2, 3).type_as(q)
        return attn.to_out[1](attn.to_out[0](o))


def _get_qkv_projections(attn, hidden_states, encoder_hidden_states):
    if encoder_hidden_states is None:
        encoder_hidden_states = hidden_states
    return attn.to_q(hidden_states), attn.to_k(encoder_hidden_states), attn.to_v(encoder_hidden_states)


def _get_added_kv_projections(attn, encoder_hidden_states_img):
    return attn.add_k_proj(encoder_hidden_states_img), attn.add_v_proj(encoder_hidden_states_img)


def dispatch_attention_fn(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, backend=None,
                          **kwargs):
    """[B, L, H, d] in and out"""
    o = F.scaled_dot_product_attention(query.transpose(1, 2), key.transpose(1, 2), value.transpose(1, 2),
                                       attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale)
    return o.transpose(1, 2)


class _GELUProj(nn.Module):
    def __init__(self, d, inner):
        super().__init__()
        self.proj = nn.Linear(d, inner)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    def __init__(self, d, inner):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(d, inner), nn.Dropout(0.0), nn.Linear(inner, d)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class WanTransformerBlock(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg["dim"]
        self.norm1 = FP32LayerNorm(d, cfg["eps"], False)
        self.attn1 = WanAttention(cfg)
        self.attn2 = WanAttention(cfg)
        self.norm2 = FP32LayerNorm(d, cfg["eps"], True)
        self.ffn = FeedForward(d, cfg["ffn_dim"])
        self.norm3 = FP32LayerNorm(d, cfg["eps"], False)
        self.scale_shift_table = nn.Parameter(torch.zeros(1, 6, d))


class _TimeEmbedding(nn.Module):
    def __init__(self, cin, d):
        super().__init__()
        self.linear_1, self.act, self.linear_2 = nn.Linear(cin, d), nn.SiLU(), nn.Linear(d, d)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class _TextProjection(nn.Module):
    def __init__(self, cin, d):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(cin, d), nn.Linear(d, d)

    def forward(self, x):
        return self.linear_2(F.gelu(self.linear_1(x), approximate="tanh"))


class _Timesteps(nn.Module):
    def __init__(self, num_channels):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, t):
        return WO.timesteps_diffusers(t, self.num_channels)


class WanTimeTextImageEmbedding(nn.Module):
    """diffusers condition embedder (T2V: no image branch): (temb, timestep_proj, text, None)"""

    def __init__(self, cfg):
        super().__init__()
        d = cfg["dim"]
        self.timesteps_proj = _Timesteps(cfg["freq_dim"])
        self.time_embedder = _TimeEmbedding(cfg["freq_dim"], d)
        self.act_fn = nn.SiLU()
        self.time_proj = nn.Linear(d, 6 * d)
        self.text_embedder = _TextProjection(cfg["text_dim"], d)

    def forward(self, timestep, encoder_hidden_states, encoder_hidden_states_image=None, timestep_seq_len=None):
        t = self.timesteps_proj(timestep)
        if timestep_seq_len is not None:
            t = t.unflatten(0, (-1, timestep_seq_len))
        dt = next(iter(self.time_embedder.parameters())).dtype
        if t.dtype != dt and dt != torch.int8:
            t = t.to(dt)
        temb = self.time_embedder(t).type_as(encoder_hidden_states)
        timestep_proj = self.time_proj(self.act_fn(temb))
        return temb, timestep_proj, self.text_embedder(encoder_hidden_states), None


class WanTransformer3DModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg["dim"]
        pt, ph, pw = cfg["patch"]
        self.config = types.SimpleNamespace(patch_size=cfg["patch"], num_attention_heads=cfg["heads"],
                                            attention_head_dim=cfg["head_dim"], num_layers=cfg["layers"])
        self.rope = WanRotaryPosEmbed(cfg)
        self.patch_embedding = nn.Conv3d(cfg["in_channels"], d, (pt, ph, pw), (pt, ph, pw))
        self.condition_embedder = WanTimeTextImageEmbedding(cfg)
        self.blocks = nn.ModuleList([WanTransformerBlock(cfg) for _ in range(cfg["layers"])])
        self.norm_out = FP32LayerNorm(d, cfg["eps"], False)
        self.proj_out = nn.Linear(d, cfg["out_channels"] * pt * ph * pw)
        self.scale_shift_table = nn.Parameter(torch.zeros(1, 2, d))
        self.gradient_checkpointing = False


# ---------------------------------------------------------------------------------------------------------
# import the reference with these stand-ins registered as `diffusers`
# ---------------------------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_wan():
    """returns (fastgen.networks.Wan.network, fastgen.networks.Wan.network_causal)"""
    if not os.path.isdir(os.path.join(REF_ROOT, "fastgen")):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    install_shims()
    d = sys.modules.get("diffusers") or _mod("diffusers")
    d.UniPCMultistepScheduler = type("UniPCMultistepScheduler", (), {})
    dm = _mod("diffusers.models", WanTransformer3DModel=WanTransformer3DModel, AutoencoderKLWan=type("AutoencoderKLWan", (), {}))
    dmt = _mod("diffusers.models.transformers")
    tw = _mod("diffusers.models.transformers.transformer_wan", WanTransformerBlock=WanTransformerBlock,
              WanRotaryPosEmbed=WanRotaryPosEmbed, WanAttnProcessor=WanAttnProcessor, WanAttention=WanAttention,
              _get_qkv_projections=_get_qkv_projections, _get_added_kv_projections=_get_added_kv_projections)
    du = _mod("diffusers.utils", USE_PEFT_BACKEND=False, scale_lora_layers=lambda *a, **k: None,
              unscale_lora_layers=lambda *a, **k: None)
    ad = _mod("diffusers.models.attention_dispatch", dispatch_attention_fn=dispatch_attention_fn)
    d.models, d.utils, dm.transformers, dm.attention_dispatch, dmt.transformer_wan = dm, du, dmt, ad, tw
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.environ.setdefault("FASTGEN_DISABLE_FLEX_ATTENTION", "1")   # dense-mask SDPA instead of compiled flex attention
    return (importlib.import_module("fastgen.networks.Wan.network"),
            importlib.import_module("fastgen.networks.Wan.network_causal"))


def build_bidirectional(cfg: Dict[str, Any], sd: Dict[str, torch.Tensor], *, r_timestep=False, time_cond_type="diff",
                        encoder_depth: Optional[int] = None, norm_temb=False, official_sinusoidal=False):
    """a stand-in WanTransformer3DModel carrying the oracle's weights, with the reference's forward functions bound exactly
    as `Wan.override_transformer_forward` binds them (Wan/network.py:836-851, 612-619)"""
    net_mod, _ = load_reference_wan()
    tr = WanTransformer3DModel(cfg)
    for blk in tr.blocks:
        blk.attn1.set_processor(WanAttnProcessor())
        blk.attn2.set_processor(WanAttnProcessor())
    tr.norm_temb, tr.encoder_depth, tr.time_cond_type = norm_temb, encoder_depth, time_cond_type
    tr.r_embedder = None
    if r_timestep:
        re = WanTimeTextImageEmbedding(cfg)
        del re.text_embedder
        tr.r_embedder = re
    inner = {k[len("transformer."):]: v for k, v in sd.items()
             if k.startswith("transformer.") and (r_timestep or ".r_embedder." not in k)}
    missing, unexpected = tr.load_state_dict(inner, strict=False)
    assert not unexpected and not [m for m in missing if "freqs" not in m], (missing, unexpected)
    for blk in tr.blocks:
        blk.forward = types.MethodType(net_mod.block_forward, blk)
    tr.classify_forward_prepare = types.MethodType(net_mod.classify_forward_prepare, tr)
    tr.classify_forward_block_forward = types.MethodType(net_mod.classify_forward_block_forward, tr)
    tr.forward = types.MethodType(net_mod.classify_forward, tr)
    if official_sinusoidal:
        for emb in (tr.condition_embedder, tr.r_embedder):
            if emb is not None:
                nc = emb.timesteps_proj.num_channels
                emb.timesteps_proj.forward = types.MethodType(
                    lambda self, t, nc=nc: net_mod.sinusoidal_embedding_1d_wan(nc, t), emb.timesteps_proj)
    return tr.eval()


def build_causal(cfg: Dict[str, Any], sd: Dict[str, torch.Tensor], *, r_timestep=False, time_cond_type="diff",
                 encoder_depth: Optional[int] = None, norm_temb=False):
    """stand-in transformer with the reference's CAUSAL functions bound exactly as `CausalWan.override_transformer_forward`
    binds them (network_causal.py:1003-1028)"""
    net_mod, c = load_reference_wan()
    WanAttnProcessor._attention_backend = None
    tr = WanTransformer3DModel(cfg)
    tr.norm_temb, tr.encoder_depth, tr.time_cond_type = norm_temb, encoder_depth, time_cond_type
    tr.r_embedder = None
    if r_timestep:
        re = WanTimeTextImageEmbedding(cfg)
        del re.text_embedder
        tr.r_embedder = re
    inner = {k[len("transformer."):]: v for k, v in sd.items()
             if k.startswith("transformer.") and (r_timestep or ".r_embedder." not in k)}
    missing, unexpected = tr.load_state_dict(inner, strict=False)
    assert not unexpected and not [m for m in missing if "freqs" not in m], (missing, unexpected)
    tr.rope.forward = types.MethodType(c._rope_forward_with_time_offset, tr.rope)
    tr._prepare_blockwise_causal_attn_mask = types.MethodType(c._prepare_blockwise_causal_attn_mask, tr)
    tr._create_external_caches = types.MethodType(c._create_external_caches, tr)
    tr.set_attn_processor = types.MethodType(c._wan_set_attn_processor, tr)
    tr.classify_forward_prepare = types.MethodType(c.classify_forward_prepare, tr)
    tr.classify_forward_block_forward = types.MethodType(c.classify_forward_block_forward, tr)
    tr.forward = types.MethodType(c.classify_forward, tr)
    for blk in tr.blocks:
        blk.forward = types.MethodType(c._wan_block_forward_inline_cache, blk)
    tr.set_attn_processor(c.CausalWanAttnProcessor())
    tr.block_mask = None
    return tr.eval()


def causal_call(tr, x, timestep, text, *, chunk_size, total_num_frames, cur_start_frame=0, store_kv=False, is_ar=False,
                cache_tag="pos", r_timestep=None):
    """the transformer call of CausalWan.forward (network_causal.py:1130-1150)"""
    kw = {"cache_tag": cache_tag, "chunk_size": chunk_size, "store_kv": store_kv, "is_ar": is_ar,
          "cur_start_frame": cur_start_frame, "total_num_frames": total_num_frames}
    return tr(hidden_states=x, timestep=timestep, encoder_hidden_states=text, r_timestep=r_timestep, attention_kwargs=kw,
              feature_indices=set())
