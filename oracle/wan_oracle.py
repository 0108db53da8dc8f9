"""CPU oracle for the Wan 2.1 transformer (bidirectional and causal / KV-cached forward) in plain PyTorch.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the
product path (fastgen_b200/ has no import of this package and raises if the CUDA library is missing).

PARITY: the REFERENCE'S OWN CODE IS PINNED, THE THIRD-PARTY MODULE INTERNALS ARE NOT.  The layer arithmetic of the
reference's Wan networks lives in the third-party dependency diffusers==0.35.1 (reference requirements.txt:6;
`WanTransformer3DModel`, `WanTransformerBlock`, `WanAttention`, `WanRotaryPosEmbed`, `WanTimeTextImageEmbedding`,
imported at fastgen/networks/Wan/network.py:33-35 and network_causal.py:31-38), which is not under /root/reference and
not installed in this image, and the reference's own tests for these networks assert shapes/dtypes only
(tests/test_network.py:617-685, 933-1057): there is no golden vector for the diffusers modules themselves.  But the
functions the reference wrote around that dependency run unmodified on stand-in modules (oracle/ref_wan_harness.py), and
tests/test_oracle_golden.py::test_wan_oracle_matches_reference_glue checks this file against their outputs
(tests/golden/wan_reference_glue.npz): bidirectional forward incl. every r-embedder / encoder_depth / norm_temb /
official-sinusoid variant (fp32 and bf16), causal KV-cached chunks (denoise and cache-fill passes, rotary frame offset),
the block-causal mask through eager FlexAttention -- bit-exact, resp. 3e-7 for FlexAttention.  Reference behaviours found
that way and NOT reproduced: (1) with FlexAttention available, a full-length `is_ar=True` call pads q/k/v with zero rows to
a multiple of 128 and attends to the zero keys because no block mask is passed (network_causal.py:414-441; 8 zero keys
next to 32 760 real ones at 480p) -- this file and the CUDA path use plain attention, which is also what the reference
does without FlexAttention; (2) the causal network's r embedder cannot run with per-frame timesteps (`timestep -
r_timestep` mixes the flattened [B*F] with [B, F], :645, and the flattened r projection is unflattened as if it were 3-D,
:657) -- implemented here as the bidirectional code does it; (3) if diffusers does not keep the rotary tables in fp32
under `.to(bfloat16)`, a bf16 reference rotates with bf16-rounded cos / sin (4e-3 on a block output); fp32 tables here.
The lines followed:

  fastgen/networks/Wan/network.py:81-103        sinusoidal_embedding_1d_wan (official Wan sinusoid, fp64)
  fastgen/networks/Wan/network.py:106-153       block_forward (AdaLN modulate / gate / residual dtype chain)
  fastgen/networks/Wan/network.py:156-279       classify_forward (head modulation, proj_out, unpatchify permute)
  fastgen/networks/Wan/network.py:282-371       classify_forward_prepare (patch embedding, condition embedder)
  fastgen/networks/Wan/network.py:904-917       _compute_timestep_inputs (rescale_t = 1000 t)
  fastgen/networks/Wan/network_causal.py:79-128     _rope_forward_with_time_offset (3-axis tables, frame offset)
  fastgen/networks/Wan/network_causal.py:131-196    _prepare_blockwise_causal_attn_mask
  fastgen/networks/Wan/network_causal.py:199-464    CausalWanAttnProcessor.__call__ (RMSNorm, RoPE, KV cache)
  fastgen/networks/Wan/network_causal.py:467-548    _wan_block_forward_inline_cache (per-frame modulation)
  fastgen/networks/Wan/network_causal.py:1058-1184  CausalWan._compute_timestep_inputs / forward

and, for the pieces inside diffusers, its published module definitions restated from the 0.35.1 sources
(transformer_wan.py, embeddings.py, normalization.py, attention.py): FP32LayerNorm, torch.nn.RMSNorm across heads
(eps 1e-6), get_1d_rotary_pos_embed(use_real, repeat_interleave_real, fp64), Timesteps(flip_sin_to_cos=True,
downscale_freq_shift=0), TimestepEmbedding(linear_1, SiLU, linear_2), PixArtAlphaTextProjection(gelu_tanh),
FeedForward(gelu-approximate), F.scaled_dot_product_attention.  State-dict key names are the diffusers ones
(`transformer.blocks.0.attn1.to_q.weight`, ...), which is what the reference checkpoints hold.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .device import sdpa

WAN_1_3B = dict(dim=1536, heads=12, head_dim=128, ffn_dim=8960, layers=30, in_channels=16, out_channels=16,
                text_dim=4096, freq_dim=256, patch=(1, 2, 2), eps=1e-6, rope_max_seq_len=1024)
WAN_14B = dict(WAN_1_3B, dim=5120, heads=40, ffn_dim=13824, layers=40)


def tiny_config(**kw) -> Dict[str, Any]:
    cfg = dict(dim=256, heads=2, head_dim=128, ffn_dim=512, layers=2, in_channels=16, out_channels=16, text_dim=64,
               freq_dim=256, patch=(1, 2, 2), eps=1e-6, rope_max_seq_len=1024)
    cfg.update(kw)
    return cfg


def make_state_dict(cfg: Dict[str, Any], seed: int, dtype: torch.dtype = torch.float32,
                    r_embedder: bool = False) -> Dict[str, torch.Tensor]:
    """Random, non-degenerate weights under the diffusers key names (proj_out is NOT zero, unlike reset_parameters)."""
    g = torch.Generator().manual_seed(seed)
    D, Fd, T = cfg["dim"], cfg["ffn_dim"], cfg["text_dim"]
    pt, ph, pw = cfg["patch"]
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, cout, cin, bias=True, std=None):
        sd[name + ".weight"] = torch.randn(cout, cin, generator=g) * (std if std is not None else cin ** -0.5)
        if bias:
            sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    p = "transformer."
    sd[p + "patch_embedding.weight"] = torch.randn(D, cfg["in_channels"], pt, ph, pw, generator=g) * (
        cfg["in_channels"] * pt * ph * pw) ** -0.5
    sd[p + "patch_embedding.bias"] = torch.randn(D, generator=g) * 0.02
    lin(p + "condition_embedder.time_embedder.linear_1", D, cfg["freq_dim"])
    lin(p + "condition_embedder.time_embedder.linear_2", D, D)
    lin(p + "condition_embedder.time_proj", 6 * D, D)
    lin(p + "condition_embedder.text_embedder.linear_1", D, T)
    lin(p + "condition_embedder.text_embedder.linear_2", D, D)
    for i in range(cfg["layers"]):
        b = f"{p}blocks.{i}."
        sd[b + "scale_shift_table"] = torch.randn(1, 6, D, generator=g) / D ** 0.5
        for a in ("attn1", "attn2"):
            for nm in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(b + f"{a}.{nm}", D, D)
            sd[b + f"{a}.norm_q.weight"] = 1 + 0.1 * torch.randn(D, generator=g)
            sd[b + f"{a}.norm_k.weight"] = 1 + 0.1 * torch.randn(D, generator=g)
        sd[b + "norm2.weight"] = 1 + 0.1 * torch.randn(D, generator=g)
        sd[b + "norm2.bias"] = 0.1 * torch.randn(D, generator=g)
        lin(b + "ffn.net.0.proj", Fd, D)
        lin(b + "ffn.net.2", D, Fd)
    sd[p + "scale_shift_table"] = torch.randn(1, 2, D, generator=g) / D ** 0.5
    lin(p + "proj_out", cfg["out_channels"] * pt * ph * pw, D)
    if r_embedder:   # drawn last: the other weights do not depend on this option
        lin(p + "r_embedder.time_embedder.linear_1", D, cfg["freq_dim"])
        lin(p + "r_embedder.time_embedder.linear_2", D, D)
        lin(p + "r_embedder.time_proj", 6 * D, D)
    return {k: v.to(dtype) for k, v in sd.items()}


# ---------------------------------------------------------------------------------------------------------
# pieces of diffusers restated
# ---------------------------------------------------------------------------------------------------------
def rope_tables(cfg: Dict[str, Any]) -> Tuple[torch.Tensor, torch.Tensor]:
    """WanRotaryPosEmbed buffers freqs_cos / freqs_sin: [max_seq_len, head_dim] fp32, (t | h | w) blocks, every
    frequency repeated twice (repeat_interleave_real)."""
    d = cfg["head_dim"]
    h_dim = w_dim = 2 * (d // 6)
    t_dim = d - h_dim - w_dim
    pos = torch.arange(cfg["rope_max_seq_len"], dtype=torch.float64)
    cos, sin = [], []
    for dim in (t_dim, h_dim, w_dim):
        freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
        ang = torch.outer(pos, freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos, 1), torch.cat(sin, 1)


def rope_for_grid(cfg, ppf: int, pph: int, ppw: int, start_frame: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """network_causal.py:79-128 (start_frame = 0 gives the bidirectional WanRotaryPosEmbed.forward)."""
    fc, fs = rope_tables(cfg)
    d = cfg["head_dim"]
    split = [d - 2 * (d // 3), d // 3, d // 3]
    ct, ch, cw = fc.split(split, dim=1)
    st, sh, sw = fs.split(split, dim=1)
    total_f = ct.shape[0]
    start = max(0, start_frame)
    end = start + ppf
    if start >= total_f:
        cos_f, sin_f = ct[-1:].expand(ppf, -1), st[-1:].expand(ppf, -1)
    else:
        end = min(end, total_f)
        cos_f, sin_f = ct[start:end], st[start:end]
        if cos_f.shape[0] < ppf:
            pad = ppf - cos_f.shape[0]
            cos_f = torch.cat([cos_f, ct[end - 1:end].expand(pad, -1)], 0)
            sin_f = torch.cat([sin_f, st[end - 1:end].expand(pad, -1)], 0)

    def grid(a_f, a_h, a_w):
        a_f = a_f.view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1)
        a_h = a_h[:pph].view(1, pph, 1, -1).expand(ppf, pph, ppw, -1)
        a_w = a_w[:ppw].view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1)
        return torch.cat([a_f, a_h, a_w], -1).reshape(1, ppf * pph * ppw, 1, -1)

    return grid(cos_f, ch, cw), grid(sin_f, sh, sw)


def apply_rotary_emb(x: torch.Tensor, freqs_cos: torch.Tensor, freqs_sin: torch.Tensor) -> torch.Tensor:
    x1, x2 = x.unflatten(-1, (-1, 2)).unbind(-1)
    cos = freqs_cos[..., 0::2]
    sin = freqs_sin[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = x1 * cos - x2 * sin
    out[..., 1::2] = x1 * sin + x2 * cos
    return out.type_as(x)


def fp32_layer_norm(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """diffusers FP32LayerNorm: computes in fp32, returns the INPUT dtype (callers pass x.float())."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float() if w is not None else None,
                        b.float() if b is not None else None, eps).to(x.dtype)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """torch.nn.RMSNorm: fp32 statistics, result in x.dtype, then the weight multiply in x.dtype."""
    xf = x.float()
    y = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x)
    return y * w.to(x.dtype)


def timesteps_diffusers(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], fp32."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def sinusoidal_embedding_1d_wan(dim: int, position: torch.Tensor) -> torch.Tensor:
    """fastgen/networks/Wan/network.py:81-103."""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1).float()


def _linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype) if name + ".bias" in sd else None)


class KVCache:
    """external self / cross KV caches of one cache tag (network_causal.py:703-814): zero-initialised buffers
    [B, capacity, H, d] per block, plus the static cross-attention K/V."""

    def __init__(self, cfg, batch: int, capacity_tokens: int, text_len: int, dtype: torch.dtype, device=None):
        h, d = cfg["heads"], cfg["head_dim"]
        self.k = [torch.zeros(batch, capacity_tokens, h, d, dtype=dtype, device=device) for _ in range(cfg["layers"])]
        self.v = [torch.zeros(batch, capacity_tokens, h, d, dtype=dtype, device=device) for _ in range(cfg["layers"])]
        self.cross: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None] * cfg["layers"]


def attention(sd, pre, cfg, hidden, enc=None, rotary=None, cache: Optional[KVCache] = None, layer: int = 0,
              cache_start: Optional[int] = None, store_kv: bool = False, mask: Optional[torch.Tensor] = None):
    """CausalWanAttnProcessor.__call__ (network_causal.py:199-464) without the I2V image branch; with cache=None and
    mask=None it is diffusers' WanAttnProcessor."""
    H = cfg["heads"]
    src = hidden if enc is None else enc
    q = _linear(sd, pre + "to_q", hidden)
    k = _linear(sd, pre + "to_k", src)
    v = _linear(sd, pre + "to_v", src)
    q = rms_norm(q, sd[pre + "norm_q.weight"], cfg["eps"])
    k = rms_norm(k, sd[pre + "norm_k.weight"], cfg["eps"])
    q, k, v = (t.unflatten(2, (H, -1)) for t in (q, k, v))
    if rotary is not None:
        q = apply_rotary_emb(q, *rotary)
        k = apply_rotary_emb(k, *rotary)
    q, k = q.type_as(hidden), k.type_as(hidden)
    if enc is not None:
        if cache is not None:
            if cache.cross[layer] is not None:      # initialised and no-grad: reuse (:337-341)
                k, v = cache.cross[layer]
            elif store_kv:
                cache.cross[layer] = (k, v)
    elif cache is not None:
        seqlen = q.shape[1]
        end = cache_start + seqlen
        if store_kv:
            cache.k[layer][:, cache_start:end] = k
            cache.v[layer][:, cache_start:end] = v
        if cache_start > 0:
            k = torch.cat([cache.k[layer][:, :cache_start], k], dim=1)
            v = torch.cat([cache.v[layer][:, :cache_start], v], dim=1)
    o = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), mask)
    o = o.transpose(1, 2).flatten(2, 3).type_as(q)
    return _linear(sd, pre + "to_out.0", o)


def block_causal_mask(num_frames: int, frame_seqlen: int, chunk_size: int) -> torch.Tensor:
    """network_causal.py:131-196 as a dense boolean [N, N] mask (True = attend), without the 128-padding."""
    n = num_frames * frame_seqlen
    ends = torch.zeros(n, dtype=torch.long)
    num_chunks, rem = num_frames // chunk_size, num_frames % chunk_size
    counts = [rem] if num_chunks == 0 else [chunk_size + rem] + [chunk_size] * (num_chunks - 1)
    cur = 0
    for c in counts:
        ends[cur:cur + c * frame_seqlen] = cur + c * frame_seqlen
        cur += c * frame_seqlen
    idx = torch.arange(n)
    return (idx[None, :] < ends[:, None]) | (idx[None, :] == idx[:, None])


def normalize(x: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """fastgen/networks/Wan/network.py:72-78 with dim=None: the norm runs over every dim but the first"""
    norm = torch.linalg.vector_norm(x, dim=list(range(1, x.ndim)), keepdim=True, dtype=torch.float32)
    return x / (eps + norm).to(x.dtype)


def block(sd, i, cfg, hs, temb, enc, rotary, per_frame: bool, cache=None, cache_start=None, store_kv=False, mask=None,
          norm_temb: bool = False):
    """block_forward (Wan/network.py:106-153) / _wan_block_forward_inline_cache (network_causal.py:467-548)."""
    b = f"transformer.blocks.{i}."
    table = sd[b + "scale_shift_table"].float()
    if per_frame:  # temb [B, F, 6, D]
        mods = (table.unsqueeze(0) + temb.float()).chunk(6, dim=2)  # each [B, F, 1, D]
        nf = mods[0].shape[1]
        fs = hs.shape[1] // nf

        def modulate(x, scale, shift):
            return (x.unflatten(1, (nf, fs)) * (1 + scale) + shift).flatten(1, 2)

        def gated(x, gate):
            return (x.unflatten(1, (nf, fs)) * gate).flatten(1, 2)
    else:          # temb [B, 6, D]
        mods = (table + temb.float()).chunk(6, dim=1)               # each [B, 1, D]

        def modulate(x, scale, shift):
            return x * (1 + scale) + shift

        def gated(x, gate):
            return x * gate
    shift_msa, scale_msa, gate_msa, c_shift, c_scale, c_gate = mods
    if norm_temb:   # Wan/network.py:131-136 (the per-frame variant normalises after squeeze(2): over frames and channels)
        sq = (lambda v: v.squeeze(2)) if per_frame else (lambda v: v)
        un = (lambda v: v.unsqueeze(2)) if per_frame else (lambda v: v)
        shift_msa, scale_msa = un(normalize(sq(shift_msa))), un(normalize(sq(scale_msa)))
        c_shift, c_scale = un(normalize(sq(c_shift))), un(normalize(sq(c_scale)))
    eps = cfg["eps"]
    n1 = modulate(fp32_layer_norm(hs.float(), None, None, eps), scale_msa, shift_msa).type_as(hs)
    a = attention(sd, b + "attn1.", cfg, n1, rotary=rotary, cache=cache, layer=i, cache_start=cache_start,
                  store_kv=store_kv, mask=mask)
    hs = (hs.float() + gated(a, gate_msa)).type_as(hs)
    n2 = fp32_layer_norm(hs.float(), sd[b + "norm2.weight"], sd[b + "norm2.bias"], eps).type_as(hs)
    a = attention(sd, b + "attn2.", cfg, n2, enc=enc, cache=cache, layer=i, store_kv=store_kv)
    hs = hs + a
    n3 = modulate(fp32_layer_norm(hs.float(), None, None, eps), c_scale, c_shift).type_as(hs)
    ff = _linear(sd, b + "ffn.net.2", F.gelu(_linear(sd, b + "ffn.net.0.proj", n3), approximate="tanh"))
    return (hs.float() + gated(ff.float(), c_gate)).type_as(hs)


def transformer_forward(sd, cfg, x: torch.Tensor, timestep: torch.Tensor, text: torch.Tensor, *,
                        official_sinusoidal: bool = False, cache: Optional[KVCache] = None, cur_start_frame: int = 0,
                        store_kv: bool = False, mask: Optional[torch.Tensor] = None,
                        r_timestep: Optional[torch.Tensor] = None, time_cond_type: str = "diff",
                        encoder_depth: Optional[int] = None, norm_temb: bool = False) -> torch.Tensor:
    """classify_forward (Wan/network.py:156-279) with either prepare / block-forward flavour; r_timestep (same shape as
    timestep, rescaled) goes through the r embedder of classify_forward_prepare (:330-353).

    x [B, C, F, H, W] in the compute dtype; timestep already rescaled (1000 t): [B] (bidirectional, global AdaLN) or
    [B, F] (causal, per-frame AdaLN); text [B, L, text_dim].  Returns the raw network output [B, C, F, H, W]."""
    dt = x.dtype
    B, C, Fr, Hh, Ww = x.shape
    pt, ph, pw = cfg["patch"]
    ppf, pph, ppw = Fr // pt, Hh // ph, Ww // pw
    D = cfg["dim"]
    p = "transformer."
    rotary = tuple(r.to(x.device) for r in rope_for_grid(cfg, ppf, pph, ppw, cur_start_frame))
    hs = F.conv3d(x, sd[p + "patch_embedding.weight"].to(dt), sd[p + "patch_embedding.bias"].to(dt), stride=(pt, ph, pw))
    hs = hs.flatten(2).transpose(1, 2)
    per_frame = timestep.ndim == 2
    ts = timestep.flatten()
    # (the small sinusoid tables are always evaluated on the CPU, so a GPU evaluation of this oracle differs from the CPU
    # one by float32 layer arithmetic only, not by a different libm)
    ts_h = ts.cpu()
    emb = (sinusoidal_embedding_1d_wan(cfg["freq_dim"], ts_h) if official_sinusoidal
           else timesteps_diffusers(ts_h, cfg["freq_dim"])).to(x.device)
    if per_frame:
        emb = emb.unflatten(0, (B, -1))
    emb = emb.to(dt)
    temb = _linear(sd, p + "condition_embedder.time_embedder.linear_2",
                   F.silu(_linear(sd, p + "condition_embedder.time_embedder.linear_1", emb))).type_as(text.to(dt))
    tproj = _linear(sd, p + "condition_embedder.time_proj", F.silu(temb))
    tproj = tproj.unflatten(2, (6, -1)) if per_frame else tproj.unflatten(1, (6, -1))
    enc = _linear(sd, p + "condition_embedder.text_embedder.linear_2",
                  F.gelu(_linear(sd, p + "condition_embedder.text_embedder.linear_1", text.to(dt)), approximate="tanh"))
    rproj = None
    if r_timestep is not None:
        if p + "r_embedder.time_proj.weight" not in sd:
            raise ValueError("r_timestep provided but no r_embedder is present")
        rt = r_timestep.flatten().cpu()
        if time_cond_type == "diff":
            rt = ts_h - rt
        remb_in = (sinusoidal_embedding_1d_wan(cfg["freq_dim"], rt) if official_sinusoidal
                   else timesteps_diffusers(rt, cfg["freq_dim"])).to(x.device)
        if per_frame:
            remb_in = remb_in.unflatten(0, (B, -1))
        remb = _linear(sd, p + "r_embedder.time_embedder.linear_2",
                       F.silu(_linear(sd, p + "r_embedder.time_embedder.linear_1", remb_in.to(dt)))).type_as(text.to(dt))
        rproj = _linear(sd, p + "r_embedder.time_proj", F.silu(remb))
        rproj = rproj.unflatten(2, (6, -1)) if per_frame else rproj.unflatten(1, (6, -1))
        if encoder_depth is None:
            tproj = tproj + rproj
            temb = temb + remb
        else:
            temb = remb
    cache_start = cur_start_frame * (pph * ppw) if cache is not None else None
    for i in range(cfg["layers"]):
        if encoder_depth is not None and i == encoder_depth and rproj is not None:
            tproj = rproj
        hs = block(sd, i, cfg, hs, tproj, enc, rotary, per_frame, cache=cache, cache_start=cache_start,
                   store_kv=store_kv, mask=mask, norm_temb=norm_temb)
    table = sd[p + "scale_shift_table"]  # head: table + temb in the parameter dtype (no .float(), Wan/network.py:229,250)
    if temb.dim() == 3:   # per frame
        shift, scale = (table.unsqueeze(0) + temb.unsqueeze(2)).chunk(2, dim=2)
        shift, scale = shift.squeeze(2), scale.squeeze(2)
        fsq = pph * ppw
        hn = fp32_layer_norm(hs.float(), None, None, cfg["eps"]).unflatten(1, (ppf, fsq))
        hs = (hn * (1 + scale.unsqueeze(2)) + shift.unsqueeze(2)).flatten(1, 2).type_as(hs)
    else:
        shift, scale = (table + temb.unsqueeze(1)).chunk(2, dim=1)
        hs = (fp32_layer_norm(hs.float(), None, None, cfg["eps"]) * (1 + scale) + shift).type_as(hs)
    hs = _linear(sd, p + "proj_out", hs)
    hs = hs.reshape(B, ppf, pph, ppw, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
    return hs.flatten(6, 7).flatten(4, 5).flatten(2, 3)


# ---------------------------------------------------------------------------------------------------------
# network-level wrappers (the FastGenNetwork.forward contract)
# ---------------------------------------------------------------------------------------------------------
def _flow_to(x_t, flow, t, target: str):
    """RF schedule conversions in fp64, one rounding (noise_schedule.py:1426-1486, 666-726)."""
    tt = t.double()
    if tt.dim() <= 1:
        tt = tt.reshape(-1, *([1] * (x_t.dim() - 1)))
    if target == "flow":
        return flow
    x0 = (x_t.double() - tt * flow.double()).to(x_t.dtype)
    if target == "x0":
        return x0
    raise NotImplementedError(target)


def wan_forward(sd, cfg, x_t, t, text, fwd_pred_type="x0", dtype=torch.bfloat16, official_sinusoidal=False, r=None,
                time_cond_type="diff", encoder_depth=None, norm_temb=False):
    """Wan.forward (Wan/network.py:1077-1163), net_pred_type='flow', schedule 'rf'."""
    timestep = (t.double() * 1000.0).to(t.dtype)            # rescale_t (noise_schedule.py:1337-1341)
    r_timestep = None if r is None else (r.double() * 1000.0).to(r.dtype)
    out = transformer_forward({k: v for k, v in sd.items()}, cfg, x_t.to(dtype), timestep, text.to(dtype),
                              official_sinusoidal=official_sinusoidal, r_timestep=r_timestep,
                              time_cond_type=time_cond_type, encoder_depth=encoder_depth, norm_temb=norm_temb)
    return _flow_to(x_t, out.to(x_t.dtype), t, fwd_pred_type)


def causal_wan_forward(sd, cfg, x_t, t, text, *, cache: Optional[KVCache] = None, cur_start_frame=0, store_kv=False,
                       total_num_frames=21, chunk_size=3, is_ar=False, fwd_pred_type="x0", dtype=torch.bfloat16,
                       official_sinusoidal=False):
    """CausalWan.forward (network_causal.py:1082-1184): per-frame timesteps; KV cache when the chunk is shorter than
    total_num_frames, the block-causal mask when it is the full sequence and not autoregressive."""
    B, _, Fr, Hh, Ww = x_t.shape
    ts = (t.double() * 1000.0).to(t.dtype)
    if ts.ndim == 1:
        ts = ts.view(-1, 1)
    timestep = torch.ones(B, Fr, dtype=ts.dtype, device=ts.device) * ts
    mask = None
    use_cache = cache if Fr < total_num_frames else None
    if use_cache is None and not is_ar:
        _, ph, pw = cfg["patch"]
        mask = block_causal_mask(Fr, (Hh // ph) * (Ww // pw), chunk_size).to(x_t.device)
    out = transformer_forward(sd, cfg, x_t.to(dtype), timestep, text.to(dtype), official_sinusoidal=official_sinusoidal,
                              cache=use_cache, cur_start_frame=cur_start_frame, store_kv=store_kv, mask=mask)
    t_conv = t[:, None, :, None, None] if t.ndim == 2 else t
    return _flow_to(x_t, out.to(x_t.dtype), t_conv, fwd_pred_type)
