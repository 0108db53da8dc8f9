"""DiT (adaLN-zero diffusion transformer) on the B200 kernel library.

Drop-in for ``fastgen.networks.DiT.network.DiT`` (reference DiT/network.py:228-574): same constructor keywords, same
``forward`` signature, same ``state_dict`` key names (``blocks.0.attention.qkv.weight`` ...) including the Facebook-DiT
checkpoint key conversion (:346-401), optional ``r`` embedder (MeanFlow-style second timestep).  Inference only (no
features / logvar outputs).

Forward = patchify -> GEMM (+pos_embed) -> per block [ln_modulate -> GEMM(qkv) -> flash attention -> GEMM(proj; gate*v +
residual) -> ln_modulate -> GEMM(fc1, GELU-tanh) -> GEMM(fc2; gate*v + residual)] -> ln_modulate -> GEMM -> unpatchify;
the conditioning vectors of ALL blocks come from one GEMM on SiLU(t_emb + y).  bf16 operands / residual stream, fp32
accumulation, statistics, modulation and softmax.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, Mapping, Optional, Set

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import ops
from .network import FastGenNetwork
from .noise_schedule import NET_PRED_TYPES
from .wan import _Plan

__all__ = ["DiT"]


def _encode_1d(positions: np.ndarray, dim: int, max_wavelength: float = 10000.0) -> np.ndarray:
    half = dim // 2
    freq = 1.0 / (max_wavelength ** (np.arange(half, dtype=np.float64) / float(half)))
    ang = np.outer(positions, freq)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def compute_sinusoidal_2d_embeddings(embedding_dim: int, spatial_size: int) -> np.ndarray:
    """reference DiT/network.py:712-757"""
    assert embedding_dim % 2 == 0, "Embedding dimension must be even"
    y, x = np.meshgrid(np.arange(spatial_size, dtype=np.float32), np.arange(spatial_size, dtype=np.float32), indexing="ij")
    return np.concatenate([_encode_1d(x.reshape(-1), embedding_dim // 2), _encode_1d(y.reshape(-1), embedding_dim // 2)], axis=1)


class _PatchEmbed(nn.Module):
    def __init__(self, cin, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, patch, patch, bias=True)


class _TimeEmbedding(nn.Module):
    def __init__(self, dim, freq=256):
        super().__init__()
        self.proj_net = nn.Sequential(nn.Linear(freq, dim), nn.SiLU(), nn.Linear(dim, dim))


class _ClassEmbedding(nn.Module):
    def __init__(self, n, dim):
        super().__init__()
        self.class_embeddings = nn.Embedding(n, dim)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.attention, self.feed_forward = _Attention(dim), _Mlp(dim, hidden)
        self.conditioning_net = nn.Sequential(nn.SiLU(), nn.Linear(dim, 6 * dim))


class _Final(nn.Module):
    def __init__(self, dim, out):
        super().__init__()
        self.projection = nn.Linear(dim, out)
        self.adaptive_params = nn.Sequential(nn.SiLU(), nn.Linear(dim, 2 * dim))


class DiT(FastGenNetwork):
    def __init__(self, input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, enable_class_dropout=False, num_classes=1000, learn_sigma=False,
                 r_timestep=False, scale_t=True, time_cond_type="abs", net_pred_type="flow", schedule_type="rf",
                 enable_fused_attn=False, use_sit_convention=False, **model_kwargs):
        super().__init__(net_pred_type=net_pred_type, schedule_type=schedule_type, **model_kwargs)
        if time_cond_type not in ("abs", "diff"):
            raise ValueError(f"Invalid time_cond_type: {time_cond_type}")
        if hidden_size % num_heads or (hidden_size // num_heads) % 8 or hidden_size // num_heads > 128:
            raise ValueError("head dim must be a multiple of 8 and <= 128")
        self.use_sit_convention = use_sit_convention
        self.input_size, self.learn_sigma, self.in_channels = input_size, learn_sigma, in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size, self.num_heads, self.num_classes = patch_size, num_heads, num_classes
        self.hidden_size, self.depth, self.scale_t = hidden_size, depth, scale_t
        self.img_resolution = input_size * 8
        self.x_embedder = _PatchEmbed(in_channels, hidden_size, patch_size)
        self.t_embedder = _TimeEmbedding(hidden_size)
        self.r_timestep, self.time_cond_type = r_timestep, time_cond_type
        self.r_embedder = _TimeEmbedding(hidden_size) if r_timestep else None      # reference DiT/network.py:274-279
        n_emb = num_classes + (1 if class_dropout_prob > 0 else 0)
        self.y_embedder = _ClassEmbedding(n_emb, hidden_size)
        self.num_patches = (input_size // patch_size) ** 2
        self.register_buffer("pos_embed", torch.zeros(1, self.num_patches, hidden_size), persistent=True)
        self.blocks = nn.ModuleList([_Block(hidden_size, int(hidden_size * mlp_ratio)) for _ in range(depth)])
        self.final_layer = _Final(hidden_size, patch_size * patch_size * self.out_channels)
        self.logvar_linear = nn.Linear(hidden_size, 1)
        self.initialize_weights()
        self._packed: Optional[Dict[str, Any]] = None
        self._plans: Dict[Any, Any] = {}
        self.profile = None

    def initialize_weights(self):
        """reference DiT/network.py:289-331"""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        grid = int(self.num_patches ** 0.5)
        self.pos_embed.data.copy_(torch.from_numpy(compute_sinusoidal_2d_embeddings(self.hidden_size, grid)).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.class_embeddings.weight, std=0.02)
        nn.init.normal_(self.t_embedder.proj_net[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.proj_net[2].weight, std=0.02)
        if self.r_embedder is not None:
            nn.init.normal_(self.r_embedder.proj_net[0].weight, std=0.02)
            nn.init.normal_(self.r_embedder.proj_net[2].weight, std=0.02)
        for blk in self.blocks:
            nn.init.constant_(blk.conditioning_net[-1].weight, 0)
            nn.init.constant_(blk.conditioning_net[-1].bias, 0)
        nn.init.constant_(self.final_layer.adaptive_params[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaptive_params[-1].bias, 0)
        nn.init.constant_(self.final_layer.projection.weight, 0)
        nn.init.constant_(self.final_layer.projection.bias, 0)

    def _invalidate(self):
        self._packed, self._plans = None, {}

    def _apply(self, fn, *a, **kw):
        self._invalidate()
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, state_dict: Mapping[str, Any], **kwargs):
        """Facebook-DiT key conversion of the reference (DiT/network.py:346-401)."""
        if any(k.startswith("t_embedder.mlp") or ".attn." in k or ".adaLN_modulation" in k for k in state_dict.keys()):
            new = {}
            for key, value in state_dict.items():
                if key == "pos_embed":
                    if value.shape == self.pos_embed.shape:
                        self.pos_embed.data.copy_(value)
                    continue
                k = key.replace("t_embedder.mlp", "t_embedder.proj_net").replace(".attn.", ".attention.")
                k = k.replace(".mlp.", ".feed_forward.").replace("y_embedder.embedding_table", "y_embedder.class_embeddings")
                k = k.replace("final_layer.linear", "final_layer.projection")
                k = k.replace("final_layer.adaLN_modulation", "final_layer.adaptive_params")
                k = k.replace(".adaLN_modulation.", ".conditioning_net.")
                new[k] = value
            state_dict = new
        self._invalidate()
        return super().load_state_dict(state_dict, **kwargs)

    @torch.no_grad()
    def repack(self) -> None:
        dev = self.final_layer.projection.weight.device
        if dev.type != "cuda":
            raise L.FastGenB200Error("DiT (B200) must live on a CUDA device: there is no CPU path")
        L.lib()
        bf, f32 = torch.bfloat16, torch.float32
        w = lambda lin: lin.weight.detach().to(bf).contiguous()        # noqa: E731
        b = lambda lin: lin.bias.detach().to(f32).contiguous()         # noqa: E731
        pk: Dict[str, Any] = {}
        pe = self.x_embedder.proj
        pk["patch.w"], pk["patch.b"] = pe.weight.detach().reshape(pe.weight.shape[0], -1).to(bf).contiguous(), b(pe)
        pk["t1.w"], pk["t1.b"] = w(self.t_embedder.proj_net[0]), b(self.t_embedder.proj_net[0])
        pk["t2.w"], pk["t2.b"] = w(self.t_embedder.proj_net[2]), b(self.t_embedder.proj_net[2])
        if self.r_embedder is not None:
            pk["r1.w"], pk["r1.b"] = w(self.r_embedder.proj_net[0]), b(self.r_embedder.proj_net[0])
            pk["r2.w"], pk["r2.b"] = w(self.r_embedder.proj_net[2]), b(self.r_embedder.proj_net[2])
        cw = [blk.conditioning_net[1].weight for blk in self.blocks] + [self.final_layer.adaptive_params[1].weight]
        cb = [blk.conditioning_net[1].bias for blk in self.blocks] + [self.final_layer.adaptive_params[1].bias]
        pk["cond.w"] = torch.cat(cw, 0).detach().to(bf).contiguous()      # [(6L + 2) D, D]
        pk["cond.b"] = torch.cat(cb, 0).detach().to(f32).contiguous()
        for i, blk in enumerate(self.blocks):
            pk[f"{i}.qkv.w"], pk[f"{i}.qkv.b"] = w(blk.attention.qkv), b(blk.attention.qkv)
            pk[f"{i}.o.w"], pk[f"{i}.o.b"] = w(blk.attention.proj), b(blk.attention.proj)
            pk[f"{i}.f1.w"], pk[f"{i}.f1.b"] = w(blk.feed_forward.fc1), b(blk.feed_forward.fc1)
            pk[f"{i}.f2.w"], pk[f"{i}.f2.b"] = w(blk.feed_forward.fc2), b(blk.feed_forward.fc2)
        pk["out.w"], pk["out.b"] = w(self.final_layer.projection), b(self.final_layer.projection)
        self._packed, self._plans = pk, {}

    def _plan(self, B: int, x_dtype: torch.dtype, has_r: bool = False):
        key = (B, x_dtype, has_r)
        ent = self._plans.get(key)
        if ent is not None:
            return ent
        pk, lib = self._packed, L.lib()
        dev = self.final_layer.projection.weight.device
        D, N, heads, Lb = self.hidden_size, self.num_patches, self.num_heads, self.depth
        d, p, Cc, Co = D // heads, self.patch_size, self.in_channels, self.out_channels
        Hd = self.blocks[0].feed_forward.fc1.out_features
        rows, S = B * N, self.input_size
        kin, kout = Cc * p * p, p * p * Co
        kin_pad = (kin + 7) // 8 * 8
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt=bf: torch.zeros(*s, device=dev, dtype=dt)  # noqa: E731
        ws = dict(x_in=z(B, Cc, 1, S, S, dt=x_dtype), tok=z(rows, kin_pad), pos=z(rows, D), hs=z(rows, D), nrm=z(rows, D),
                  qkv=z(rows, 3 * D), att=z(rows, D), f1=z(rows, Hd), tok_out=z(rows, kout),
                  out=z(B, Co, 1, S, S, dt=x_dtype), emb=z(B, 256), te1=z(B, D), yemb=z(B, D, dt=f32), c=z(B, D, dt=f32),
                  emb_r=z(B, 256), re1=z(B, D), c0=z(B, D, dt=f32),
                  c_act=z(B, D), mod=z(B, (6 * Lb + 2) * D, dt=f32))
        ws["pos"].copy_(self.pos_embed.detach().to(bf).expand(B, N, D).reshape(rows, D))
        if kin_pad != kin:
            pw = z(D, kin_pad)
            pw[:, :kin] = pk["patch.w"]
            patch_w = pw
        else:
            patch_w = pk["patch.w"]
        probe = L.FlashAttnParams()
        probe.batch, probe.heads, probe.d, probe.tq = B, heads, d, N
        ws["attn_ws"] = torch.empty(max(ops.flash_attn_workspace_size(probe), 16), device=dev, dtype=torch.uint8)
        plan = _Plan()
        plan.keep.append((ws, patch_w))

        def gemm(a, k, wt, r, cout, out, out_ld, desc, **kw):
            prm = ops.conv_gemm_params([ops.Seg(a, a.stride(0), wt, k, 1)], r, 1, 1, cout, out, out_ld, **kw)
            fn, args, pkeep = ops.conv_gemm_launch(prm)
            plan.keep.append((prm, pkeep, a, wt, out, kw))
            plan.add(fn, *args, desc=desc, flops=2.0 * r * k * cout)

        def ln(x, y, scale, shift, desc):
            prm = ops.ln_mod_params(x, D, rows, D, 1e-6, y, D, scale=scale, shift=shift, mod_ld=ws["mod"].shape[1],
                                    rows_per_group=N)
            plan.keep.append(prm)
            plan.add(lib.fg_ln_modulate, C.byref(prm), desc=desc)

        mod_ld = ws["mod"].shape[1]
        plan.add(lib.fg_patchify, ws["x_in"].data_ptr(), ops._DT[x_dtype], B, Cc, 1, S, S, p, p, ws["tok"].data_ptr(),
                 kin_pad, desc="patchify")
        gemm(ws["tok"], kin_pad, patch_w, rows, D, ws["hs"], D, "patch", bias=pk["patch.b"], res=ws["pos"], res_ld=D,
             res_mode=L.FG_RES_SAME)
        gemm(ws["emb"], 256, pk["t1.w"], B, D, ws["te1"], D, "time1", bias=pk["t1.b"], act=L.FG_ACT_SILU)
        if not has_r:
            gemm(ws["te1"], D, pk["t2.w"], B, D, ws["c"], D, "time2", bias=pk["t2.b"], res=ws["yemb"], res_ld=D,
                 res_mode=L.FG_RES_SAME)                                # c = t_emb + y
        else:                                                           # c = t_emb + y + r_emb (DiT/network.py:518-533)
            gemm(ws["te1"], D, pk["t2.w"], B, D, ws["c0"], D, "time2", bias=pk["t2.b"], res=ws["yemb"], res_ld=D,
                 res_mode=L.FG_RES_SAME)
            gemm(ws["emb_r"], 256, pk["r1.w"], B, D, ws["re1"], D, "r_time1", bias=pk["r1.b"], act=L.FG_ACT_SILU)
            gemm(ws["re1"], D, pk["r2.w"], B, D, ws["c"], D, "r_time2", bias=pk["r2.b"], res=ws["c0"], res_ld=D,
                 res_mode=L.FG_RES_SAME)
        plan.add(lib.fg_split_bf16, ws["c"].data_ptr(), D, None, 0, B, D, 1, ws["c_act"].data_ptr(), None, D, None, 0,
                 desc="silu_c")
        gemm(ws["c_act"], D, pk["cond.w"], B, mod_ld, ws["mod"], mod_ld, "cond", bias=pk["cond.b"])
        hs, nrm, qkv, att, f1 = ws["hs"], ws["nrm"], ws["qkv"], ws["att"], ws["f1"]
        for i in range(Lb):
            m = lambda j: ws["mod"][:, (6 * i + j) * D:(6 * i + j + 1) * D]   # noqa: E731
            ln(hs, nrm, m(1), m(0), "ln1")
            gemm(nrm, D, pk[f"{i}.qkv.w"], rows, 3 * D, qkv, 3 * D, "qkv", bias=pk[f"{i}.qkv.b"])
            prm = ops.flash_attn_params(qkv, 3 * D, N * 3 * D, [(qkv[:, D:], qkv[:, 2 * D:], 3 * D, N * 3 * D, N)], B, heads,
                                        d, N, att, D, N * D, workspace=ws["attn_ws"])
            plan.keep.append(prm)
            plan.add(lib.fg_flash_attn, C.byref(prm), desc="attn", flops=4.0 * B * N * N * D)
            gemm(att, D, pk[f"{i}.o.w"], rows, D, hs, D, "proj", bias=pk[f"{i}.o.b"], gate=m(2), gate_ld=mod_ld, gate_rows=N,
                 res=hs, res_ld=D, res_mode=L.FG_RES_SAME)
            ln(hs, nrm, m(4), m(3), "ln2")
            gemm(nrm, D, pk[f"{i}.f1.w"], rows, Hd, f1, Hd, "fc1", bias=pk[f"{i}.f1.b"], act=L.FG_ACT_GELU_TANH)
            gemm(f1, Hd, pk[f"{i}.f2.w"], rows, D, hs, D, "fc2", bias=pk[f"{i}.f2.b"], gate=m(5), gate_ld=mod_ld,
                 gate_rows=N, res=hs, res_ld=D, res_mode=L.FG_RES_SAME)
        fm = ws["mod"][:, 6 * Lb * D:]
        ln(hs, nrm, fm[:, D:], fm[:, :D], "ln_out")
        gemm(nrm, D, pk["out.w"], rows, kout, ws["tok_out"], kout, "proj_out", bias=pk["out.b"])
        plan.add(lib.fg_unpatchify, ws["tok_out"].data_ptr(), kout, B, Co, 1, S, S, p, p, ws["out"].data_ptr(),
                 ops._DT[x_dtype], desc="unpatchify")
        self._plans[key] = (plan, ws)
        return plan, ws

    def prepare_t(self, t, dtype=None):
        if t is None:
            return t
        if self.scale_t:
            t = self.noise_scheduler.rescale_t(t)
        return t.to(dtype=dtype)

    def forward(self, x_t: torch.Tensor, t: torch.Tensor, condition: Optional[torch.Tensor] = None,
                r: Optional[torch.Tensor] = None, return_features_early: bool = False,
                feature_indices: Optional[Set[int]] = None, return_logvar: bool = False,
                fwd_pred_type: Optional[str] = None, **fwd_kwargs):
        if feature_indices is None:
            feature_indices = {}
        if fwd_pred_type is None:
            fwd_pred_type = self.net_pred_type
        else:
            assert fwd_pred_type in NET_PRED_TYPES, f"{fwd_pred_type} is not supported as fwd_pred_type"
        if return_features_early or len(feature_indices) or return_logvar:
            raise NotImplementedError("feature / logvar outputs are training-time paths of the reference")
        if not x_t.is_cuda:
            raise L.FastGenB200Error("DiT (B200) needs CUDA tensors: there is no CPU path")
        if x_t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"x_t must be float32 or bfloat16, got {x_t.dtype}")
        with torch.no_grad():
            if self._packed is None:
                self.repack()
            B = x_t.shape[0]
            if tuple(x_t.shape[1:]) != (self.in_channels, self.input_size, self.input_size):
                raise ValueError(f"x_t shape {tuple(x_t.shape)} does not match the network")
            x_in, t_in = x_t, t
            if condition.ndim == 2:
                mask = torch.any(condition != 0, dim=1)
                condition = torch.where(~mask, self.num_classes, condition.argmax(dim=1))
            tt = self.prepare_t(torch.atleast_1d(t).to(x_t.device), dtype=x_t.dtype)
            if tt.numel() == 1 and B > 1:
                tt = tt.expand(B)
            if self.use_sit_convention:
                tt = 1 - tt
            has_r = self.r_embedder is not None and r is not None     # (r without an r embedder is ignored, like the reference)
            plan, ws = self._plan(B, x_t.dtype, has_r)
            ws["x_in"].copy_(x_t.reshape(ws["x_in"].shape))
            half = 128
            freq = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=x_t.device) / half)

            def fourier(v):
                ang = v[:, None].float() * freq[None, :]
                return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(v.dtype)

            ws["emb"].copy_(fourier(tt))
            if has_r:
                rr = self.prepare_t(torch.atleast_1d(r).to(x_t.device), dtype=x_t.dtype)
                if rr.numel() == 1 and B > 1:
                    rr = rr.expand(B)
                ws["emb_r"].copy_(fourier(tt - rr if self.time_cond_type == "diff" else rr))
            ws["yemb"].copy_(self.y_embedder.class_embeddings.weight.detach()[condition.to(x_t.device)])
            plan.run(torch.cuda.current_stream().cuda_stream, self.profile)
            out = ws["out"].reshape(B, self.out_channels, self.input_size, self.input_size)
            if self.learn_sigma:
                out = out.chunk(2, dim=1)[0]
            out = out.clone()
            if self.use_sit_convention and self.net_pred_type == "flow":
                out = -out
            return self.noise_scheduler.convert_model_output(x_in, out, torch.atleast_1d(t_in).to(x_t.device),
                                                             src_pred_type=self.net_pred_type,
                                                             target_pred_type=fwd_pred_type)
