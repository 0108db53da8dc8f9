"""Wan 2.1 T2V transformers (bidirectional ``Wan`` and chunk-autoregressive ``CausalWan``) on the B200 kernel library.

Drop-ins for ``fastgen.networks.Wan.network.Wan`` (reference Wan/network.py:540-1163) and
``fastgen.networks.Wan.network_causal.CausalWan`` (network_causal.py:928-1184): same ``forward`` signature and keyword
semantics (``cache_tag / cur_start_frame / store_kv / is_ar``, unknown keywords ignored), same ``state_dict`` key names
as the diffusers ``WanTransformer3DModel`` the reference wraps (``transformer.blocks.0.attn1.to_q.weight`` ...), the
original-Wan checkpoint key remap of Wan/network.py:990-1075, ``clear_caches()``, ``chunk_size``,
``total_num_frames``, ``noise_scheduler``.  Inference only, T2V only (no I2V image branch, no ``r`` embedder).

How a forward runs (PyTorch only allocates buffers and computes a [B*F, 256] sinusoid):
  patchify -> GEMM (Conv3d as a K=64 GEMM) -> per block:
      ln_modulate -> GEMM(qkv) -> qk_norm_rope (RMSNorm across heads + RoPE, in place, + KV-cache append)
      -> flash_attn(keys = cache prefix ++ current chunk) -> GEMM(to_out; gate * v + residual)
      -> ln_modulate(norm2 affine) -> GEMM(q) -> qk_norm -> flash_attn(512 cached text keys) -> GEMM(to_out + residual)
      -> ln_modulate -> GEMM(ffn up, GELU-tanh) -> GEMM(ffn down; gate * v + residual)
  -> ln_modulate(head) -> GEMM(proj_out) -> unpatchify -> flow -> x0 conversion (fp64, one rounding).
The residual stream and every GEMM operand are bf16 (the reference runs these networks in bf16), accumulators,
LayerNorm / RMSNorm statistics, modulation vectors and softmax are fp32.  The launch list of a (shape, cache position)
is recorded once as pre-filled C-ABI parameter structs and replayed.

Sequence parallelism (Ulysses): with ``sp_group`` set, every rank holds tokens [rank * n/P, (rank+1) * n/P) of each
sample, q/k/v are exchanged by an all-to-all (scatter heads, gather tokens) before each self-attention and the
attention output by the inverse exchange; the KV cache is head-sharded, so cached keys never move again.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Any, Dict, List, Mapping, Optional, Set, Tuple

import torch
from torch import nn

from . import _lib as L
from . import ops
from . import parallel as PL
from .network import CausalFastGenNetwork, FastGenNetwork, tensor_identity
from .noise_schedule import NET_PRED_TYPES

__all__ = ["Wan", "CausalWan", "WAN_CONFIGS"]

WAN_CONFIGS: Dict[str, Dict[str, Any]] = {
    "Wan-AI/Wan2.1-T2V-1.3B-Diffusers": dict(dim=1536, heads=12, head_dim=128, ffn_dim=8960, layers=30),
    "Wan-AI/Wan2.1-T2V-14B-Diffusers": dict(dim=5120, heads=40, head_dim=128, ffn_dim=13824, layers=40),
}
_DEFAULTS = dict(in_channels=16, out_channels=16, text_dim=4096, freq_dim=256, patch=(1, 2, 2), eps=1e-6,
                 rope_max_seq_len=1024)
TEXT_LEN = 512
# tests/test_wan_host_cpu.py sets this to build launch lists against host memory, so that the argument validation of
# every C-ABI call can be exercised on a CPU-only host; such a plan cannot execute (every launch fails).
_ALLOW_HOST_PLANS = False


# ---------------------------------------------------------------------------------------------------------
# parameter tree with the diffusers key names
# ---------------------------------------------------------------------------------------------------------
class _Attn(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])
        self.norm_q = nn.RMSNorm(dim, eps=1e-6)
        self.norm_k = nn.RMSNorm(dim, eps=1e-6)


class _GeluProj(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner)


class _FFN(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.net = nn.ModuleList([_GeluProj(dim, inner), nn.Identity(), nn.Linear(inner, dim)])


class _Block(nn.Module):
    def __init__(self, dim: int, ffn_dim: int):
        super().__init__()
        self.scale_shift_table = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)
        self.attn1, self.attn2 = _Attn(dim), _Attn(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6, elementwise_affine=True)
        self.ffn = _FFN(dim, ffn_dim)


class _TimeEmbedder(nn.Module):
    def __init__(self, cin: int, dim: int):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(cin, dim), nn.Linear(dim, dim)


class _CondEmbedder(nn.Module):
    def __init__(self, dim: int, freq_dim: int, text_dim: int):
        super().__init__()
        self.time_embedder = _TimeEmbedder(freq_dim, dim)
        self.time_proj = nn.Linear(dim, 6 * dim)
        self.text_embedder = _TimeEmbedder(text_dim, dim)


class _REmbedder(nn.Module):
    """the reference's r embedder: a copy of condition_embedder without its text branch (Wan/network.py:799-834)"""

    def __init__(self, dim: int, freq_dim: int):
        super().__init__()
        self.time_embedder = _TimeEmbedder(freq_dim, dim)
        self.time_proj = nn.Linear(dim, 6 * dim)


class _WanTransformer(nn.Module):
    def __init__(self, cfg: Dict[str, Any], logvar: bool):
        super().__init__()
        d = cfg["dim"]
        pt, ph, pw = cfg["patch"]
        self.patch_embedding = nn.Conv3d(cfg["in_channels"], d, (pt, ph, pw), (pt, ph, pw))
        self.condition_embedder = _CondEmbedder(d, cfg["freq_dim"], cfg["text_dim"])
        self.blocks = nn.ModuleList([_Block(d, cfg["ffn_dim"]) for _ in range(cfg["layers"])])
        self.scale_shift_table = nn.Parameter(torch.randn(1, 2, d) / d ** 0.5)
        self.proj_out = nn.Linear(d, cfg["out_channels"] * pt * ph * pw)
        if logvar:
            self.logvar_linear = nn.Linear(d, 1)


def _rope_pair_tables(head_dim: int, max_pos: int, theta: float = 10000.0):
    """WanRotaryPosEmbed reduced to one entry per rotated pair: cos/sin [max_pos, head_dim/2] fp32 with the temporal,
    height and width blocks side by side (== freqs_cos[..., 0::2], freqs_sin[..., 1::2])."""
    h_dim = w_dim = 2 * (head_dim // 6)
    t_dim = head_dim - h_dim - w_dim
    pos = torch.arange(max_pos, dtype=torch.float64)
    cos, sin = [], []
    for dim in (t_dim, h_dim, w_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
        ang = torch.outer(pos, freqs)
        cos.append(ang.cos().float())
        sin.append(ang.sin().float())
    return torch.cat(cos, 1).contiguous(), torch.cat(sin, 1).contiguous(), t_dim // 2, h_dim // 2


def _sinusoid(t: torch.Tensor, dim: int, official: bool) -> torch.Tensor:
    """[len(t), dim] = [cos | sin]: the official Wan sinusoid in fp64 (reference Wan/network.py:81-103) or diffusers'
    ``Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0)`` in fp32."""
    half = dim // 2
    if official:
        pos = t.to(torch.float64)
        s = torch.outer(pos, torch.pow(10000, -torch.arange(half, device=t.device).to(pos).div(half)))
        return torch.cat([torch.cos(s), torch.sin(s)], dim=1).float()
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


# ---------------------------------------------------------------------------------------------------------
# recorded launch list
# ---------------------------------------------------------------------------------------------------------
class _Plan:
    def __init__(self):
        self.steps: List[Tuple[Any, tuple, str, float]] = []
        self.keep: List[Any] = []
        self.runs = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.n_head = 0
        self.has_r = False
        self.mod_groups = (0, 0)

    def add(self, fn, *args, desc: str = "", flops: float = 0.0):
        self.steps.append((fn, args, desc, flops))

    def add_py(self, fn, desc: str):
        """a host-side step (a torch.distributed collective on the current stream)"""
        self.steps.append((fn, None, desc, 0.0))

    def run(self, stream: int, profile: Optional[Dict[str, Dict[str, float]]] = None) -> None:
        if profile is None:
            for fn, args, _, _ in self.steps:
                if args is None:
                    fn()
                    continue
                st = fn(*args, stream)
                if st:
                    L.check(st, fn.__name__)
            return
        evs = []
        for fn, args, desc, fl in self.steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if args is None:
                fn()
                st = 0
            else:
                st = fn(*args, stream)
            b.record()
            if st:
                L.check(st, fn.__name__)
            evs.append(((fn.__name__ if args is not None else "collective") + (":" + desc if desc else ""), a, b, fl))
        torch.cuda.synchronize()
        for name, a, b, fl in evs:
            d = profile.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0})
            d["launches"] += 1
            d["ms"] += a.elapsed_time(b)
            d["flops"] += fl


class _WanBase:
    """Shared implementation (mixed into the two public classes)."""

    def _init_wan(self, model_id_or_local_path, transformer_config, enable_logvar_linear, use_wan_official_sinusoidal,
                  r_timestep, norm_temb, encoder_depth, r_embedder_init="pretrained", time_cond_type="diff"):
        if time_cond_type not in ("abs", "diff"):
            raise ValueError(f"Invalid time condition: {time_cond_type}")
        if r_embedder_init not in ("pretrained", "zero", "random"):
            raise ValueError(f"Invalid embedder_init: {r_embedder_init}")
        cfg = dict(_DEFAULTS)
        if transformer_config is not None:
            cfg.update(transformer_config)
        else:
            model_id = self.get_model_id(model_id_or_local_path)
            if model_id not in WAN_CONFIGS:
                raise ValueError(f"unknown Wan model id '{model_id}' (known: {sorted(WAN_CONFIGS)}); "
                                 "pass transformer_config=... for a custom architecture")
            cfg.update(WAN_CONFIGS[model_id])
        if cfg["dim"] != cfg["heads"] * cfg["head_dim"]:
            raise ValueError("dim must equal heads * head_dim")
        if cfg["patch"][0] != 1:
            raise NotImplementedError("temporal patch size must be 1")
        self.cfg = cfg
        self.model_id = model_id_or_local_path
        self.expand_timesteps = False
        self._use_wan_official_sinusoidal = use_wan_official_sinusoidal
        self.transformer = _WanTransformer(cfg, enable_logvar_linear)
        self.norm_temb, self.encoder_depth, self.time_cond_type = bool(norm_temb), encoder_depth, time_cond_type
        if encoder_depth is not None and not 0 <= encoder_depth <= cfg["layers"]:
            raise ValueError("encoder_depth must be in [0, layers]")
        self.reset_wan_parameters()
        self.transformer.r_embedder = self._init_r_embedder(r_embedder_init) if r_timestep else None
        self._packed: Optional[Dict[str, Any]] = None
        self._plans: Dict[Any, _Plan] = {}
        self._ws: Dict[Any, Dict[str, torch.Tensor]] = {}
        self._kv: Dict[Any, Dict[str, Any]] = {}      # cache_tag -> self-attention KV cache
        self._cross: Dict[Any, Dict[str, Any]] = {}   # cache_tag -> static cross-attention K/V of the current prompt
        self.profile: Optional[Dict[str, Dict[str, float]]] = None
        # replay each recorded launch list as a CUDA graph from its second use on (the kernels are short enough for
        # the ~10 us per launch of the host to show, above all under sequence parallelism)
        self.use_cuda_graphs = True
        self.graph_launches = 0                        # kernel launches replayed through graphs (bench bookkeeping)
        self.sp_group = None                           # torch.distributed group for Ulysses sequence parallelism
        # how the sequence-parallel exchange around the self-attention runs: "peer" (default) = fg_peer_put / fg_peer_wait
        # over symmetric memory: pack and exchange are one kernel storing into the peers' buffers over NVLink, no NCCL call
        # in the layer loop; "nccl" = pack kernel + all_to_all_single + unpack kernel.  Measured on 2 x B200, Wan 1.3B
        # self-forcing sample: nccl 787 ms, peer 745 ms, peer + CUDA graphs 720 ms (profiles/wan_scaling_r02.md).  If the
        # symmetric-memory rendezvous is unavailable the workspace falls back to "nccl" with a warning.
        self.sp_exchange = os.environ.get("FASTGEN_B200_SP_EXCHANGE", "peer")
        if self.sp_exchange not in ("nccl", "peer"):
            raise ValueError(f"FASTGEN_B200_SP_EXCHANGE must be 'nccl' or 'peer', got {self.sp_exchange}")
        self.peer_timeout_s = 60.0
        # with the peer exchange no NCCL call is left in a launch list, so it replays as a CUDA graph under sequence
        # parallelism too (36 k launches per sample)
        self.sp_graphs = os.environ.get("FASTGEN_B200_SP_GRAPHS", "1") == "1"

    # ---- reference helpers ------------------------------------------------------------------------------
    @classmethod
    def get_model_id(cls, model_path_or_id: str) -> str:
        """reference Wan/network.py:784-797 (a local directory name carries the model id)"""
        import os

        if not os.path.isdir(model_path_or_id):
            return model_path_or_id
        idx = model_path_or_id.rfind("Wan2.")
        if idx == -1:
            raise ValueError("unable to extract model id from path. Expects substring `Wan2.X-`")
        return "Wan-AI/" + model_path_or_id[idx:].split("/")[0]

    def reset_wan_parameters(self) -> None:
        """initialisation of reference Wan.reset_parameters (Wan/network.py:709-727)"""
        for m in self.transformer.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.transformer.patch_embedding.weight.flatten(1))
        for m in self.transformer.condition_embedder.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.02)
        nn.init.zeros_(self.transformer.proj_out.weight)

    def _init_r_embedder(self, how: str) -> "_REmbedder":
        """reference Wan.init_embedder (Wan/network.py:799-834)"""
        ce = self.transformer.condition_embedder
        emb = _REmbedder(self.cfg["dim"], self.cfg["freq_dim"]).to(ce.time_proj.weight.device, ce.time_proj.weight.dtype)
        with torch.no_grad():
            emb.time_embedder.load_state_dict(ce.time_embedder.state_dict())
            emb.time_proj.load_state_dict(ce.time_proj.state_dict())
            if how == "zero":
                for prm in emb.parameters():
                    prm.zero_()
            elif how == "random":
                for m in emb.modules():
                    if isinstance(m, nn.Linear):
                        nn.init.xavier_uniform_(m.weight)
                        nn.init.zeros_(m.bias)
                for m in emb.time_embedder.modules():
                    if isinstance(m, nn.Linear):
                        nn.init.normal_(m.weight, std=0.02)
                nn.init.zeros_(emb.time_embedder.linear_2.weight)
                nn.init.zeros_(emb.time_embedder.linear_2.bias)
                nn.init.zeros_(emb.time_proj.weight)
                nn.init.zeros_(emb.time_proj.bias)
        return emb

    def _invalidate(self):
        self._packed = None
        self._plans = {}
        self._ws = {}
        self._kv = {}
        self._cross = {}

    def _apply(self, fn, *a, **kw):
        self._invalidate()
        return super()._apply(fn, *a, **kw)

    _RENAME = [
        ("time_embedding.0", "condition_embedder.time_embedder.linear_1"),
        ("time_embedding.2", "condition_embedder.time_embedder.linear_2"),
        ("text_embedding.0", "condition_embedder.text_embedder.linear_1"),
        ("text_embedding.2", "condition_embedder.text_embedder.linear_2"),
        ("time_projection.1", "condition_embedder.time_proj"),
        ("head.modulation", "scale_shift_table"), ("head.head", "proj_out"), ("modulation", "scale_shift_table"),
        ("ffn.0", "ffn.net.0.proj"), ("ffn.2", "ffn.net.2"),
        ("norm2", "norm__placeholder"), ("norm3", "norm2"), ("norm__placeholder", "norm3"),
        ("self_attn.q", "attn1.to_q"), ("self_attn.k", "attn1.to_k"), ("self_attn.v", "attn1.to_v"),
        ("self_attn.o", "attn1.to_out.0"), ("self_attn.norm_q", "attn1.norm_q"), ("self_attn.norm_k", "attn1.norm_k"),
        ("cross_attn.q", "attn2.to_q"), ("cross_attn.k", "attn2.to_k"), ("cross_attn.v", "attn2.to_v"),
        ("cross_attn.o", "attn2.to_out.0"), ("cross_attn.norm_q", "attn2.norm_q"), ("cross_attn.norm_k", "attn2.norm_k"),
    ]

    def load_state_dict(self, state_dict: Mapping[str, Any], **kwargs):
        """Also accepts original-Wan checkpoints, optionally nested under 'generator' / 'state_dict' / ... (reference
        Wan/network.py:990-1075)."""
        if self._use_wan_official_sinusoidal and not any(k.startswith("transformer.") for k in state_dict.keys()):
            state = None
            for k in ["generator", "state_dict", "model", "module", "net", None]:
                if k is None:
                    if all(isinstance(v, torch.Tensor) for v in state_dict.values()):
                        state = state_dict
                    break
                if k in state_dict and isinstance(state_dict[k], dict):
                    state = state_dict[k]
                    break
            assert state is not None, "Could not find a state_dict in checkpoint."

            def rename(key: str) -> str:
                for prefix in ["model.", "module.", "transformer."]:
                    if key.startswith(prefix):
                        key = key[len(prefix):]
                for old, new in self._RENAME:
                    if old in key:
                        key = key.replace(old, new)
                return "transformer." + key

            state_dict = {rename(k): v for k, v in state.items()}
        self._invalidate()
        return super().load_state_dict(state_dict, **kwargs)

    # ---- weight packing ---------------------------------------------------------------------------------
    @torch.no_grad()
    def repack(self) -> None:
        tr = self.transformer
        dev = tr.proj_out.weight.device
        if dev.type != "cuda" and not _ALLOW_HOST_PLANS:
            raise L.FastGenB200Error("Wan (B200) must live on a CUDA device: there is no CPU path")
        L.lib()
        bf, f32 = torch.bfloat16, torch.float32

        def w(lin):
            return lin.weight.detach().to(bf).contiguous()

        def b(lin):
            return lin.bias.detach().to(bf).to(f32).contiguous()  # the reference adds the bf16 bias

        pk: Dict[str, Any] = {}
        pe = tr.patch_embedding
        pk["patch.w"] = pe.weight.detach().reshape(pe.weight.shape[0], -1).to(bf).contiguous()
        pk["patch.b"] = pe.bias.detach().to(bf).to(f32).contiguous()
        ce = tr.condition_embedder
        for nm, lin in (("t1", ce.time_embedder.linear_1), ("t2", ce.time_embedder.linear_2), ("tp", ce.time_proj),
                        ("x1", ce.text_embedder.linear_1), ("x2", ce.text_embedder.linear_2), ("out", tr.proj_out)):
            pk[nm + ".w"], pk[nm + ".b"] = w(lin), b(lin)
        if tr.r_embedder is not None:
            re = tr.r_embedder
            for nm, lin in (("r1", re.time_embedder.linear_1), ("r2", re.time_embedder.linear_2), ("rp", re.time_proj)):
                pk[nm + ".w"], pk[nm + ".b"] = w(lin), b(lin)
        tables = []
        for i, blk in enumerate(tr.blocks):
            a1, a2 = blk.attn1, blk.attn2
            pk[f"{i}.qkv.w"] = torch.cat([a1.to_q.weight, a1.to_k.weight, a1.to_v.weight], 0).detach().to(bf).contiguous()
            pk[f"{i}.qkv.b"] = torch.cat([a1.to_q.bias, a1.to_k.bias, a1.to_v.bias], 0).detach().to(bf).to(f32).contiguous()
            pk[f"{i}.o1.w"], pk[f"{i}.o1.b"] = w(a1.to_out[0]), b(a1.to_out[0])
            pk[f"{i}.nq1"], pk[f"{i}.nk1"] = a1.norm_q.weight.detach().to(bf).contiguous(), a1.norm_k.weight.detach().to(bf).contiguous()
            pk[f"{i}.q2.w"], pk[f"{i}.q2.b"] = w(a2.to_q), b(a2.to_q)
            pk[f"{i}.kv2.w"] = torch.cat([a2.to_k.weight, a2.to_v.weight], 0).detach().to(bf).contiguous()
            pk[f"{i}.kv2.b"] = torch.cat([a2.to_k.bias, a2.to_v.bias], 0).detach().to(bf).to(f32).contiguous()
            pk[f"{i}.o2.w"], pk[f"{i}.o2.b"] = w(a2.to_out[0]), b(a2.to_out[0])
            pk[f"{i}.nq2"], pk[f"{i}.nk2"] = a2.norm_q.weight.detach().to(bf).contiguous(), a2.norm_k.weight.detach().to(bf).contiguous()
            pk[f"{i}.n2.g"] = blk.norm2.weight.detach().to(bf).to(f32).contiguous()
            pk[f"{i}.n2.b"] = blk.norm2.bias.detach().to(bf).to(f32).contiguous()
            pk[f"{i}.f1.w"], pk[f"{i}.f1.b"] = w(blk.ffn.net[0].proj), b(blk.ffn.net[0].proj)
            pk[f"{i}.f2.w"], pk[f"{i}.f2.b"] = w(blk.ffn.net[2]), b(blk.ffn.net[2])
            tables.append(blk.scale_shift_table.detach().to(bf).to(f32).reshape(-1))
        pk["tables"] = torch.stack(tables, 0).contiguous()                       # [L, 6D] fp32
        pk["head_table"] = tr.scale_shift_table.detach().to(bf).reshape(2, -1)   # bf16: the head adds in bf16
        cos, sin, nt, nh = _rope_pair_tables(self.cfg["head_dim"], self.cfg["rope_max_seq_len"])
        pk["rope"] = (cos.to(dev), sin.to(dev), nt, nh)
        self._packed = pk
        self._plans, self._ws = {}, {}

    # ---- workspace --------------------------------------------------------------------------------------
    def _workspace(self, B: int, F: int, Hh: int, Ww: int, x_dtype: torch.dtype) -> Dict[str, torch.Tensor]:
        key = (B, F, Hh, Ww, x_dtype)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        cfg = self.cfg
        dev = self.transformer.proj_out.weight.device
        D, Fd = cfg["dim"], cfg["ffn_dim"]
        _, ph, pw = cfg["patch"]
        n = F * (Hh // ph) * (Ww // pw)
        rows = B * n
        P = self._sp_size()
        fs_tok = (Hh // ph) * (Ww // pw)
        if fs_tok % P or cfg["heads"] % P:
            raise ValueError(f"sequence parallelism over {P} ranks needs tokens per frame ({fs_tok}) and heads "
                             f"({cfg['heads']}) divisible by {P}")
        lrows = rows // P                      # rows this rank owns
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt=bf: torch.zeros(*s, device=dev, dtype=dt)  # noqa: E731
        kin = cfg["in_channels"] * ph * pw
        kout = cfg["out_channels"] * ph * pw
        groups = B * F
        Hl = cfg["heads"] // P
        Dl = Hl * cfg["head_dim"]
        ws = dict(
            x_in=z(B, cfg["in_channels"], F, Hh, Ww, dt=x_dtype), tok=z(rows, kin), hs=z(lrows, D), nrm=z(lrows, D),
            qkv=z(lrows, 3 * D), att=z(lrows, D), q2=z(lrows, D), f1=z(lrows, Fd), tok_out=z(rows, kout),
            flow=z(B, cfg["out_channels"], F, Hh, Ww, dt=x_dtype),
            emb=z(groups, cfg["freq_dim"]), te1=z(groups, D), temb=z(groups, D), temb_act=z(groups, D),
            tproj=z(groups, 6 * D, dt=f32), mod=z(cfg["layers"], groups, 6 * D, dt=f32), head_mod=z(groups, 2 * D, dt=f32),
        )
        if self.transformer.r_embedder is not None:
            ws.update(emb_r=z(groups, cfg["freq_dim"]), re1=z(groups, D), remb=z(groups, D), remb_act=z(groups, D),
                      rtproj=z(groups, 6 * D, dt=f32))
        if P > 1 and self.sp_exchange not in ("nccl", "peer"):
            raise ValueError(f"sp_exchange must be 'nccl' or 'peer', got {self.sp_exchange}")
        if P > 1 and self.sp_exchange == "peer":
            # one-pass exchange: qkv_full and att are the DESTINATIONS the peers store into (symmetric memory)
            try:
                peer = self._peer_buffers(B, n, D, Dl, P, dev, kout)
                ws.update(hs_full=z(rows, D), att_full=z(B, n, Dl), tok_out_l=z(lrows, kout))
                ws.update(peer)
            except L.FastGenB200Error:
                raise
            except Exception as e:   # no symmetric memory on this platform / group: every rank sees the same failure
                import warnings

                warnings.warn(f"sequence-parallel peer exchange unavailable ({e!r}); using the NCCL all-to-all")
                self.sp_exchange = "nccl"
                self._plans = {}
        if P > 1 and self.sp_exchange == "nccl":
            # Ulysses exchange buffers (per sample b: [P][n/P][3][H/P*d] for q,k,v and [P][n/P][H/P*d] for the output)
            ws.update(hs_full=z(rows, D), a2a_s=z(B, P, n // P, 3 * Dl), qkv_full=z(B, n, 3 * Dl), att_full=z(B, n, Dl),
                      a2a_r=z(B, P, n // P, Dl), tok_out_l=z(lrows, kout), tok_out_g=z(P, lrows, kout))
        # scratch for the split-KV attention path (partial results of up to 8 key splits)
        probe = L.FlashAttnParams()
        probe.batch, probe.heads, probe.d, probe.tq = B, cfg["heads"] // P, cfg["head_dim"], n
        ws["attn_ws"] = torch.empty(max(ops.flash_attn_workspace_size(probe), 16), device=dev, dtype=torch.uint8)
        self._ws[key] = ws
        return ws

    def _peer_buffers(self, B: int, n: int, D: int, Dl: int, P: int, dev, kout: int) -> Dict[str, Any]:
        """Symmetric-memory destinations of the one-pass exchange: [ qkv_full [B, n, 3 Dl] | att [B n/P, D] | tok_out_g
        [P, B n/P, kout] | flags of the three channels ], the same layout on every rank, so a peer's region is its mapped
        base pointer + the same offset."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        al = lambda v: (v + 255) // 256 * 256  # noqa: E731
        full_b, att_b = B * n * 3 * Dl * 2, B * (n // P) * D * 2
        gat_b = P * B * (n // P) * kout * 2
        off_att = al(full_b)
        off_gat = al(off_att + att_b)
        off_flags = al(off_gat + gat_b)
        total = off_flags + al(3 * L.FG_MAX_PEERS * 8)
        import warnings

        enable = getattr(symm, "enable_symm_mem_for_group", None)
        if enable is not None:       # registers the group's store with the allocator (a no-op where rendezvous does it)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                enable(self.sp_group.group_name)
        buf = symm.empty(total, dtype=torch.uint8, device=dev)
        buf.zero_()
        hdl = symm.rendezvous(buf, self.sp_group)
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.sp_group)      # every rank's flags are zero before anybody's first put
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != P or ptrs[self._sp_rank()] != buf.data_ptr():
            raise L.FastGenB200Error("symmetric memory rendezvous returned an unexpected pointer table")
        bf = torch.bfloat16
        return dict(
            qkv_full=buf[:full_b].view(bf).view(B, n, 3 * Dl), att=buf[off_att:off_att + att_b].view(bf).view(B * (n // P), D),
            tok_out_g=buf[off_gat:off_gat + gat_b].view(bf).view(P, B * (n // P), kout),
            peer=dict(buf=buf, hdl=hdl, full_ptrs=ptrs, att_ptrs=[p + off_att for p in ptrs],
                      gat_ptrs=[p + off_gat for p in ptrs],
                      flag_ptrs=[[p + off_flags + c * L.FG_MAX_PEERS * 8 for p in ptrs] for c in range(3)],
                      state=torch.zeros(3, 4, dtype=torch.int64, device=dev)))

    def check_peer_exchange(self) -> None:
        """Synchronise and raise if a wait of the one-pass exchange ever timed out (a peer stopped making progress)."""
        for ws in self._ws.values():
            px = ws.get("peer")
            if px is not None and bool(px["state"][:, 3].any().item()):
                raise L.FastGenB200Error(f"sequence-parallel peer exchange timed out after {self.peer_timeout_s} s: a rank "
                                         "of the group did not reach the matching put")

    def _sp_size(self) -> int:
        if self.sp_group is None:
            return 1
        import torch.distributed as dist

        return dist.get_world_size(self.sp_group)

    def _sp_rank(self) -> int:
        if self.sp_group is None:
            return 0
        import torch.distributed as dist

        return dist.get_rank(self.sp_group)

    # ---- launch helpers ---------------------------------------------------------------------------------
    def _gemm(self, plan: _Plan, a: torch.Tensor, k: int, wt: torch.Tensor, rows: int, cout: int, out: torch.Tensor,
              out_ld: int, desc: str, **kw):
        p = ops.conv_gemm_params([ops.Seg(a, a.stride(0), wt, k, 1)], rows, 1, 1, cout, out, out_ld, **kw)
        fn, args, keep = ops.conv_gemm_launch(p)
        plan.keep.append((p, keep, a, wt, out, kw))
        plan.add(fn, *args, desc=desc, flops=2.0 * rows * k * cout)

    def _ln(self, plan: _Plan, x, rows, y, desc, **kw):
        D = self.cfg["dim"]
        p = ops.ln_mod_params(x, D, rows, D, self.cfg["eps"], y, D, **kw)
        plan.keep.append((p, x, y, kw))
        plan.add(L.lib().fg_ln_modulate, C.byref(p), desc=desc)

    # ---- cross-attention K/V of a prompt (static: computed once per condition tensor) --------------------
    def _cross_kv(self, condition: torch.Tensor, cache_tag: str) -> Dict[str, Any]:
        """text embedder + per-layer k/v projections + RMSNorm(k): [L, B*Ln, 2D] bf16.  The buffers of a cache tag are
        persistent (forward plans point into them); their contents are recomputed when the condition tensor changes
        or after clear_caches() -- the reference recomputes them per call until the first store_kv pass
        (network_causal.py:326-345), with identical values."""
        key = tensor_identity(condition)
        cfg, pk = self.cfg, self._packed
        D, Ln, Td = cfg["dim"], condition.shape[1], cfg["text_dim"]
        B = condition.shape[0]
        ent = self._cross.get(cache_tag)
        if ent is None or ent["batch"] != B or ent["len"] != Ln:
            dev = condition.device
            bf = torch.bfloat16
            ent = {"key": None, "len": Ln, "batch": B,
                   "text": torch.empty(B * Ln, Td, device=dev, dtype=bf), "x1": torch.empty(B * Ln, D, device=dev, dtype=bf),
                   "enc": torch.empty(B * Ln, D, device=dev, dtype=bf),
                   "kv": torch.empty(cfg["layers"], B * Ln, 2 * D, device=dev, dtype=bf)}
            plan = _Plan()
            self._gemm(plan, ent["text"], Td, pk["x1.w"], B * Ln, D, ent["x1"], D, "text1", bias=pk["x1.b"],
                       act=L.FG_ACT_GELU_TANH)
            self._gemm(plan, ent["x1"], D, pk["x2.w"], B * Ln, D, ent["enc"], D, "text2", bias=pk["x2.b"])
            for i in range(cfg["layers"]):
                self._gemm(plan, ent["enc"], D, pk[f"{i}.kv2.w"], B * Ln, 2 * D, ent["kv"][i], 2 * D, "kv2",
                           bias=pk[f"{i}.kv2.b"])
                p = ops.qk_norm_rope_params(None, ent["kv"][i], 0, 2 * D, B * Ln, cfg["heads"], cfg["head_dim"],
                                            cfg["eps"], None, pk[f"{i}.nk2"], Ln)
                plan.keep.append(p)
                plan.add(L.lib().fg_qk_norm_rope, C.byref(p), desc="k2norm")
            ent["plan"] = plan
            self._cross[cache_tag] = ent
            self._plans = {}   # forward plans hold pointers into the previous buffers
        if key is None or ent["key"] != key:
            ent["text"].copy_(condition.reshape(B * Ln, Td))
            ent["plan"].run(torch.cuda.current_stream().cuda_stream, self.profile)
            ent["key"] = key
            ent["keep"] = condition
        return ent

    # ---- self-attention KV cache -------------------------------------------------------------------------
    def _self_kv(self, cache_tag: str, B: int, capacity: int) -> Dict[str, Any]:
        ent = self._kv.get(cache_tag)
        cfg = self.cfg
        Dl = cfg["dim"] // self._sp_size()   # head-sharded under sequence parallelism
        if ent is None or ent["k"].shape[1] != B or ent["k"].shape[2] != capacity:
            dev = self.transformer.proj_out.weight.device
            ent = {"k": torch.zeros(cfg["layers"], B, capacity, Dl, device=dev, dtype=torch.bfloat16),
                   "v": torch.zeros(cfg["layers"], B, capacity, Dl, device=dev, dtype=torch.bfloat16), "len": 0}
            self._kv[cache_tag] = ent
            self._plans = {}   # forward plans hold pointers into the previous buffers
        return ent

    def clear_caches(self) -> None:
        """reference network_causal.py:1030-1056: lengths to 0, buffers zeroed (here: in place, no re-allocation)"""
        if self.sp_group is not None and self.sp_exchange == "peer":
            self.check_peer_exchange()     # a timed-out device-side wait must not pass silently (synchronises)
        for ent in self._kv.values():
            ent["k"].zero_()
            ent["v"].zero_()
            ent["len"] = 0
        for ent in self._cross.values():
            ent["key"] = None

    # ---- the forward plan --------------------------------------------------------------------------------
    def _copy(self, plan: _Plan, src, dst, n, ss, ds, item, desc):
        args, keep = ops.strided_copy_args(src, dst, n, ss, ds, item)
        plan.keep.append(keep)
        plan.add(L.lib().fg_strided_copy, *args, desc=desc)

    def _copy_desc(self, plan: _Plan, src, dst, dsc, desc):
        self._copy(plan, src.reshape(-1)[dsc["src_off"]:], dst.reshape(-1)[dsc["dst_off"]:], dsc["n"], dsc["ss"],
                   dsc["ds"], dsc["item"], desc)

    def _build_plan(self, B, F, Hh, Ww, x_dtype, per_frame: bool, kv: Optional[Dict[str, Any]], cache_start: int,
                    store_kv: bool, frame0: int, mask: Optional[Tuple[int, int]], cross: Dict[str, Any],
                    has_r: bool = False) -> _Plan:
        """Launch list of one forward.  With sequence parallelism (P ranks) a rank owns tokens
        [rank * fs/P, (rank+1) * fs/P) of EVERY frame (so modulation groups stay contiguous), all heads; around each
        self-attention q/k/v are exchanged to (all tokens, H/P heads) and the output back (Ulysses); the KV cache holds
        this rank's heads only."""
        cfg, pk = self.cfg, self._packed
        lib = L.lib()
        D, Fd, H, d = cfg["dim"], cfg["ffn_dim"], cfg["heads"], cfg["head_dim"]
        _, ph, pw = cfg["patch"]
        pph, ppw = Hh // ph, Ww // pw
        fs = pph * ppw
        n = F * fs
        rows = B * n
        ws = self._workspace(B, F, Hh, Ww, x_dtype)
        plan = _Plan()
        plan.keep.append((ws, cross["kv"], kv))
        groups = B * F
        kin = cfg["in_channels"] * ph * pw
        kout = cfg["out_channels"] * ph * pw
        P, rank = self._sp_size(), self._sp_rank()
        fs_l, n_l = fs // P, n // P
        lrows = B * n_l                 # token rows this rank owns
        Hl, Dl = H // P, (H // P) * d   # heads this rank attends with
        if P > 1 and mask:
            raise NotImplementedError("the block-causal mask (full-sequence causal forward) is not sequence-parallel")
        # (global AdaLN under sequence parallelism: one group per sample, and the local rows of a sample are contiguous,
        #  so the modulation-group arithmetic below is the same as on one GPU)
        # ---- embeddings (steps 0..6: see _run_split) ----
        plan.add(lib.fg_patchify, ws["x_in"].data_ptr(), ops._DT[x_dtype], B, cfg["in_channels"], F, Hh, Ww, ph, pw,
                 ws["tok"].data_ptr(), kin, desc="patchify")
        self._gemm(plan, ws["tok"], kin, pk["patch.w"], rows, D, ws["hs_full"] if P > 1 else ws["hs"], D, "patch",
                   bias=pk["patch.b"])
        self._gemm(plan, ws["emb"], cfg["freq_dim"], pk["t1.w"], groups, D, ws["te1"], D, "time1", bias=pk["t1.b"],
                   act=L.FG_ACT_SILU)
        self._gemm(plan, ws["te1"], D, pk["t2.w"], groups, D, ws["temb"], D, "time2", bias=pk["t2.b"])
        self._gemm(plan, ws["te1"], D, pk["t2.w"], groups, D, ws["temb_act"], D, "time2s", bias=pk["t2.b"],
                   act=L.FG_ACT_SILU)
        self._gemm(plan, ws["temb_act"], D, pk["tp.w"], groups, 6 * D, ws["tproj"], 6 * D, "tproj", bias=pk["tp.b"])
        nl = cfg["layers"]
        if not has_r:
            plan.add(lib.fg_bcast_add, pk["tables"].data_ptr(), ws["tproj"].data_ptr(), nl, groups, 6 * D,
                     ws["mod"].data_ptr(), desc="mod")
        else:
            # r embedder (MeanFlow-style second timestep, Wan/network.py:330-352): same three linears on the embedding
            # of r (or t - r); its projection is added to the time projection, or replaces it from block
            # `encoder_depth` on
            self._gemm(plan, ws["emb_r"], cfg["freq_dim"], pk["r1.w"], groups, D, ws["re1"], D, "r_time1", bias=pk["r1.b"],
                       act=L.FG_ACT_SILU)
            self._gemm(plan, ws["re1"], D, pk["r2.w"], groups, D, ws["remb"], D, "r_time2", bias=pk["r2.b"])
            self._gemm(plan, ws["re1"], D, pk["r2.w"], groups, D, ws["remb_act"], D, "r_time2s", bias=pk["r2.b"],
                       act=L.FG_ACT_SILU)
            k = self.encoder_depth
            if k is None:
                self._gemm(plan, ws["remb_act"], D, pk["rp.w"], groups, 6 * D, ws["rtproj"], 6 * D, "r_tproj",
                           bias=pk["rp.b"], res=ws["tproj"], res_ld=6 * D, res_mode=L.FG_RES_SAME)
                plan.add(lib.fg_bcast_add, pk["tables"].data_ptr(), ws["rtproj"].data_ptr(), nl, groups, 6 * D,
                         ws["mod"].data_ptr(), desc="mod")
            else:
                self._gemm(plan, ws["remb_act"], D, pk["rp.w"], groups, 6 * D, ws["rtproj"], 6 * D, "r_tproj",
                           bias=pk["rp.b"])
                if k > 0:
                    plan.add(lib.fg_bcast_add, pk["tables"].data_ptr(), ws["tproj"].data_ptr(), k, groups, 6 * D,
                             ws["mod"].data_ptr(), desc="mod")
                if k < nl:
                    plan.add(lib.fg_bcast_add, pk["tables"][k:].data_ptr(), ws["rtproj"].data_ptr(), nl - k, groups, 6 * D,
                             ws["mod"][k:].data_ptr(), desc="mod_r")
        plan.n_head = len(plan.steps)           # embedding launches: _run_split runs them, then the host-side tail
        plan.has_r = has_r
        plan.mod_groups = (B, F if per_frame else 1)
        hs, nrm, qkv, att, q2, f1 = ws["hs"], ws["nrm"], ws["qkv"], ws["att"], ws["q2"], ws["f1"]
        if P > 1:
            # this rank's token rows (b, f, i) <- global rows (b, f, rank*fs_l + i)
            self._copy_desc(plan, ws["hs_full"], hs, PL.shard_rows_desc(B, F, fs, P, rank, D), "shard")
            import torch.distributed as dist

            group = self.sp_group
        rpg = fs_l if per_frame else n_l   # local rows sharing one modulation vector
        ckv, clen = cross["kv"], cross["len"]
        cap = kv["k"].shape[2] if kv is not None else 0
        for i in range(cfg["layers"]):
            mod = ws["mod"][i]                                      # [groups, 6D]
            m = lambda j: mod[:, j * D:(j + 1) * D]                 # noqa: E731
            # 1. self-attention
            self._ln(plan, hs, lrows, nrm, "ln1", scale=m(1), shift=m(0), mod_ld=6 * D, rows_per_group=rpg)
            self._gemm(plan, nrm, D, pk[f"{i}.qkv.w"], lrows, 3 * D, qkv, 3 * D, "qkv", bias=pk[f"{i}.qkv.b"])
            kw = {}
            if P == 1 and kv is not None and store_kv:
                kw = dict(v=qkv[:, 2 * D:], v_ld=3 * D, k_cache=kv["k"][i], v_cache=kv["v"][i], cache_ld=D,
                          cache_bs=cap * D, cache_row0=cache_start)
            p = ops.qk_norm_rope_params(qkv, qkv[:, D:], 3 * D, 3 * D, lrows, H, d, cfg["eps"], pk[f"{i}.nq1"],
                                        pk[f"{i}.nk1"], n_l, rope=pk["rope"], grid=(F, pph, ppw), frame0=frame0,
                                        frame_rows=fs_l, tok0=rank * fs_l, **kw)
            plan.keep.append(p)
            plan.add(lib.fg_qk_norm_rope, C.byref(p), desc="qknorm_rope")
            mk = dict(mask_first=mask[0], mask_chunk=mask[1]) if mask else {}
            kv_tot = n + (cache_start if kv is not None else 0)
            if P == 1:
                cur = (qkv[:, D:], qkv[:, 2 * D:], 3 * D, n * 3 * D, n)
                segs = [(kv["k"][i], kv["v"][i], D, cap * D, cache_start), cur] if kv is not None and cache_start > 0 else [cur]
                p = ops.flash_attn_params(qkv, 3 * D, n * 3 * D, segs, B, H, d, n, att, D, n * D, workspace=ws["attn_ws"], **mk)
                plan.keep.append(p)
                plan.add(lib.fg_flash_attn, C.byref(p), desc="self_attn", flops=4.0 * B * n * kv_tot * D)
            elif self.sp_exchange == "peer":
                px, full, attf = ws["peer"], ws["qkv_full"], ws["att_full"]
                tmo = int(self.peer_timeout_s * 1e9)
                # qkv[b][r][j][dst][Dl] -> rank dst's full[b][rank*n_l + r][j][Dl], then wait for everybody's slices
                # (put and wait are one launch: the CTA that publishes this rank's flags waits for the peers')
                p = ops.peer_put_params(qkv, PL.peer_put_qkv_desc(B, n, P, rank, D, Dl), px["full_ptrs"], px["flag_ptrs"][0],
                                        rank, px["state"][0], wait=True, timeout_ns=tmo)
                plan.keep.append(p)
                plan.add(lib.fg_peer_put, C.byref(p), desc="peer_put_qkv")
                if kv is not None and store_kv:
                    self._copy(plan, full[:, :, Dl:], kv["k"][i][:, cache_start:], [B, n, 1, 1], [n * 3 * Dl, 3 * Dl, 0, 0],
                               [cap * Dl, Dl, 0, 0], Dl, "k_append")
                    self._copy(plan, full[:, :, 2 * Dl:], kv["v"][i][:, cache_start:], [B, n, 1, 1],
                               [n * 3 * Dl, 3 * Dl, 0, 0], [cap * Dl, Dl, 0, 0], Dl, "v_append")
                cur = (full[:, :, Dl:], full[:, :, 2 * Dl:], 3 * Dl, n * 3 * Dl, n)
                segs = [(kv["k"][i], kv["v"][i], Dl, cap * Dl, cache_start), cur] if kv is not None and cache_start > 0 else [cur]
                p = ops.flash_attn_params(full, 3 * Dl, n * 3 * Dl, segs, B, Hl, d, n, attf, Dl, n * Dl,
                                          workspace=ws["attn_ws"])
                plan.keep.append(p)
                plan.add(lib.fg_flash_attn, C.byref(p), desc="self_attn", flops=4.0 * B * n * kv_tot * Dl)
                # att_full[b][dst*n_l + r][Dl] -> rank dst's att[b][r][rank*Dl ...]: exchange + unpack in one pass
                p = ops.peer_put_params(attf, PL.peer_put_out_desc(B, n, P, rank, D, Dl), px["att_ptrs"], px["flag_ptrs"][1],
                                        rank, px["state"][1], wait=True, timeout_ns=tmo)
                plan.keep.append(p)
                plan.add(lib.fg_peer_put, C.byref(p), desc="peer_put_out")
            else:
                send, full, attf, recv = ws["a2a_s"], ws["qkv_full"], ws["att_full"], ws["a2a_r"]
                # pack: qkv[b][r][j][dst][Dl] -> send[b][dst][r][j][Dl]   (one launch per sample keeps it 4-D)
                for b in range(B):
                    self._copy_desc(plan, qkv[b * n_l:(b + 1) * n_l], send[b], PL.a2a_pack_desc(n_l, P, D, Dl), "a2a_pack")
                    plan.add_py(lambda o=full[b], s_=send[b]: dist.all_to_all_single(o, s_, group=group), "a2a_qkv")
                # full[b] = [src][r][3][Dl] = all n tokens of the chunk (gathered order) x this rank's heads
                if kv is not None and store_kv:
                    self._copy(plan, full[:, :, Dl:], kv["k"][i][:, cache_start:], [B, n, 1, 1], [n * 3 * Dl, 3 * Dl, 0, 0],
                               [cap * Dl, Dl, 0, 0], Dl, "k_append")
                    self._copy(plan, full[:, :, 2 * Dl:], kv["v"][i][:, cache_start:], [B, n, 1, 1],
                               [n * 3 * Dl, 3 * Dl, 0, 0], [cap * Dl, Dl, 0, 0], Dl, "v_append")
                cur = (full[:, :, Dl:], full[:, :, 2 * Dl:], 3 * Dl, n * 3 * Dl, n)
                segs = [(kv["k"][i], kv["v"][i], Dl, cap * Dl, cache_start), cur] if kv is not None and cache_start > 0 else [cur]
                p = ops.flash_attn_params(full, 3 * Dl, n * 3 * Dl, segs, B, Hl, d, n, attf, Dl, n * Dl,
                                          workspace=ws["attn_ws"])
                plan.keep.append(p)
                plan.add(lib.fg_flash_attn, C.byref(p), desc="self_attn", flops=4.0 * B * n * kv_tot * Dl)
                for b in range(B):
                    plan.add_py(lambda o=recv[b], s_=attf[b]: dist.all_to_all_single(o, s_, group=group), "a2a_out")
                # unpack: recv[b][src][r][Dl] -> att[b][r][src][Dl]
                self._copy_desc(plan, recv, att, PL.a2a_unpack_desc(B, P, n_l, D, Dl), "a2a_unpack")
            self._gemm(plan, att, D, pk[f"{i}.o1.w"], lrows, D, hs, D, "o1", bias=pk[f"{i}.o1.b"], gate=m(2), gate_ld=6 * D,
                       gate_rows=rpg, res=hs, res_ld=D, res_mode=L.FG_RES_SAME)
            # 2. cross-attention (text keys replicated on every rank: no exchange)
            self._ln(plan, hs, lrows, nrm, "ln2", gamma=pk[f"{i}.n2.g"], beta=pk[f"{i}.n2.b"])
            self._gemm(plan, nrm, D, pk[f"{i}.q2.w"], lrows, D, q2, D, "q2", bias=pk[f"{i}.q2.b"])
            p = ops.qk_norm_rope_params(q2, None, D, 0, lrows, H, d, cfg["eps"], pk[f"{i}.nq2"], None, n_l)
            plan.keep.append(p)
            plan.add(lib.fg_qk_norm_rope, C.byref(p), desc="q2norm")
            p = ops.flash_attn_params(q2, D, n_l * D, [(ckv[i], ckv[i][:, D:], 2 * D, clen * 2 * D, clen)], B, H, d, n_l,
                                      att, D, n_l * D, workspace=ws["attn_ws"] if P == 1 else None)
            plan.keep.append(p)
            plan.add(lib.fg_flash_attn, C.byref(p), desc="cross_attn", flops=4.0 * B * n_l * clen * D)
            self._gemm(plan, att, D, pk[f"{i}.o2.w"], lrows, D, hs, D, "o2", bias=pk[f"{i}.o2.b"], res=hs, res_ld=D,
                       res_mode=L.FG_RES_SAME)
            # 3. feed-forward
            self._ln(plan, hs, lrows, nrm, "ln3", scale=m(4), shift=m(3), mod_ld=6 * D, rows_per_group=rpg)
            self._gemm(plan, nrm, D, pk[f"{i}.f1.w"], lrows, Fd, f1, Fd, "ffn1", bias=pk[f"{i}.f1.b"],
                       act=L.FG_ACT_GELU_TANH)
            self._gemm(plan, f1, Fd, pk[f"{i}.f2.w"], lrows, D, hs, D, "ffn2", bias=pk[f"{i}.f2.b"], gate=m(5),
                       gate_ld=6 * D, gate_rows=rpg, res=hs, res_ld=D, res_mode=L.FG_RES_SAME)
        # ---- head ----
        hm = ws["head_mod"]
        self._ln(plan, hs, lrows, nrm, "ln_out", scale=hm[:, D:], shift=hm[:, :D], mod_ld=2 * D, rows_per_group=rpg)
        if P == 1:
            self._gemm(plan, nrm, D, pk["out.w"], rows, kout, ws["tok_out"], kout, "proj_out", bias=pk["out.b"])
        else:
            self._gemm(plan, nrm, D, pk["out.w"], lrows, kout, ws["tok_out_l"], kout, "proj_out", bias=pk["out.b"])
            if self.sp_exchange == "peer":
                # all-gather as a put of this rank's rows into slot `rank` of every rank's tok_out_g (third channel)
                px = ws["peer"]
                p = ops.peer_put_params(ws["tok_out_l"], PL.peer_gather_desc(lrows, kout, rank), px["gat_ptrs"],
                                        px["flag_ptrs"][2], rank, px["state"][2], wait=True,
                                        timeout_ns=int(self.peer_timeout_s * 1e9))
                plan.keep.append(p)
                plan.add(lib.fg_peer_put, C.byref(p), desc="peer_put_gather")
            else:
                plan.add_py(lambda: dist.all_gather_into_tensor(ws["tok_out_g"], ws["tok_out_l"], group=group), "gather_out")
            # gathered[src][(b, f)][i] -> natural token order [(b, f)][src * fs_l + i]
            self._copy_desc(plan, ws["tok_out_g"], ws["tok_out"], PL.unshard_desc(P, B, F, fs, kout), "unshard")
        plan.add(lib.fg_unpatchify, ws["tok_out"].data_ptr(), kout, B, cfg["out_channels"], F, Hh, Ww, ph, pw,
                 ws["flow"].data_ptr(), ops._DT[x_dtype], desc="unpatchify")
        return plan

    def _run_transformer(self, x_t: torch.Tensor, timestep: torch.Tensor, condition: torch.Tensor, per_frame: bool,
                         cache_tag: str = "pos", use_cache: bool = False, cur_start_frame: int = 0,
                         store_kv: bool = False, total_num_frames: int = 21,
                         mask: Optional[Tuple[int, int]] = None, r_timestep: Optional[torch.Tensor] = None) -> torch.Tensor:
        """timestep: rescaled (1000 t), [B] (global AdaLN) or [B, F] (per frame).  Returns the raw output [B,C,F,H,W]."""
        if not x_t.is_cuda:
            raise L.FastGenB200Error("Wan (B200) needs CUDA tensors: there is no CPU path")
        if x_t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"x_t must be float32 or bfloat16, got {x_t.dtype}")
        if self._packed is None:
            self.repack()
        cfg = self.cfg
        B, Cc, F, Hh, Ww = x_t.shape
        _, ph, pw = cfg["patch"]
        if Cc != cfg["in_channels"] or Hh % ph or Ww % pw:
            raise ValueError(f"x_t shape {tuple(x_t.shape)} does not match the network")
        if condition is None:
            raise ValueError("Wan needs text embeddings as the condition")
        if condition.shape[0] == 1 and B > 1:
            condition = condition.expand(B, -1, -1)
        if condition.shape[0] != B or condition.shape[-1] != cfg["text_dim"]:
            raise ValueError(f"condition shape {tuple(condition.shape)} does not match batch {B} / text_dim {cfg['text_dim']}")
        fs = (Hh // ph) * (Ww // pw)
        cross = self._cross_kv(condition, cache_tag)
        kv, cache_start, frame0 = None, 0, cur_start_frame
        if use_cache:
            kv = self._self_kv(cache_tag, B, fs * total_num_frames)
            cache_start = cur_start_frame * fs
            if cache_start + F * fs > kv["k"].shape[2]:
                raise ValueError("chunk does not fit the KV cache (cur_start_frame + frames > total_num_frames)")
            if cur_start_frame == 0:
                # reference network_causal.py:605-617: without an explicit start frame the rotary offset continues
                # from the cached length (0 right after clear_caches())
                frame0 = kv["len"] // fs
            if store_kv:
                kv["len"] = cache_start + F * fs
        has_r = r_timestep is not None
        if has_r and self.transformer.r_embedder is None:
            raise ValueError("r_timestep provided but no r_embedder is present")
        key = (B, F, Hh, Ww, x_t.dtype, per_frame, use_cache, cache_start, store_kv, frame0, mask, cache_tag, has_r)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._build_plan(B, F, Hh, Ww, x_t.dtype, per_frame, kv, cache_start, store_kv, frame0, mask, cross,
                                    has_r=has_r)
            if len(self._plans) > 64:
                self._plans.clear()
            self._plans[key] = plan
        ws = self._workspace(B, F, Hh, Ww, x_t.dtype)
        # ---- stage the dynamic inputs ----
        ws["x_in"].copy_(x_t)
        ts = timestep.reshape(B, -1)
        groups = B * F
        tsf = (ts.expand(B, F) if ts.shape[1] != F else ts).reshape(-1) if per_frame else ts[:, 0]
        emb = _sinusoid(tsf, cfg["freq_dim"], self._use_wan_official_sinusoidal).to(torch.bfloat16)
        if per_frame:
            ws["emb"].copy_(emb)
        else:
            ws["emb"].zero_()
            ws["emb"][:B].copy_(emb)   # rows [0, B): one modulation group per sample (rows_per_group = n)
        if has_r:
            rs = r_timestep.reshape(B, -1)
            rsf = (rs.expand(B, F) if rs.shape[1] != F else rs).reshape(-1) if per_frame else rs[:, 0]
            if self.time_cond_type == "diff":
                rsf = tsf - rsf
            emb_r = _sinusoid(rsf, cfg["freq_dim"], self._use_wan_official_sinusoidal).to(torch.bfloat16)
            if per_frame:
                ws["emb_r"].copy_(emb_r)
            else:
                ws["emb_r"].zero_()
                ws["emb_r"][:B].copy_(emb_r)
        # the head adds table + temb in bf16 (reference Wan/network.py:229,250): needs temb, so run the embedding
        # launches first, then that one tiny add, then the rest of the plan
        # (not under sequence parallelism: capturing the NCCL all-to-all into a graph deadlocked the NCCL watchdog)
        if self.use_cuda_graphs and self.profile is None and (self.sp_group is None or (self.sp_graphs and self.sp_exchange == "peer")):
            if plan.graph is None and plan.runs >= 1:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_split(plan, ws, torch.cuda.current_stream().cuda_stream)
                plan.graph = g
            if plan.graph is not None:
                plan.graph.replay()
                self.graph_launches += sum(1 for st in plan.steps if st[1] is not None)
                plan.runs += 1
                self._poll_peer_exchange(ws)
                return ws["flow"]
        self._run_split(plan, ws, torch.cuda.current_stream().cuda_stream)
        plan.runs += 1
        self._poll_peer_exchange(ws)
        return ws["flow"]

    def _poll_peer_exchange(self, ws) -> None:
        """One-pass exchange only: queue an asynchronous copy of the device-side timeout flags behind this forward and
        look at the copy queued behind the PREVIOUS one (complete by now, or skipped without waiting).  A wait that gave up
        -- the consumer then read stale data -- therefore raises at the next forward, not only at clear_caches() /
        the end of sample(); no synchronisation is added."""
        px = ws.get("peer")
        if px is None:
            return
        ev = px.get("poll_event")
        if ev is not None and ev.query() and bool(px["poll_host"].any()):
            raise L.FastGenB200Error(f"sequence-parallel peer exchange timed out after {self.peer_timeout_s} s: a rank of "
                                     "the group did not reach the matching put (outputs since then are invalid)")
        if "poll_host" not in px:
            px["poll_host"] = torch.zeros(3, dtype=torch.int64).pin_memory()
        px["poll_host"].copy_(px["state"][:, 3], non_blocking=True)
        px["poll_event"] = torch.cuda.Event()
        px["poll_event"].record()

    def _run_split(self, plan: _Plan, ws, stream: int) -> None:
        D = self.cfg["dim"]
        pk = self._packed
        # the first plan.n_head steps are the embedding launches (patchify, patch, time1, time2, time2s, tproj[, r...], mod)
        head = _Plan()
        head.steps, rest = plan.steps[:plan.n_head], plan.steps[plan.n_head:]
        head.run(stream, self.profile)
        # the head adds table + temb in bf16 (reference Wan/network.py:229,250); with an r embedder temb is temb + remb,
        # or remb alone when encoder_depth is set (:349-353)
        temb = ws["temb"]
        if plan.has_r:
            temb = ws["remb"] if self.encoder_depth is not None else ws["temb"] + ws["remb"]
        hm = (pk["head_table"][None] + temb[:, None, :]).reshape(temb.shape[0], 2 * D)
        ws["head_mod"].copy_(hm)
        if self.norm_temb:
            # sCM-style normalisation of the AdaLN shift / scale vectors (Wan/network.py:72-78, 132-136): x / (1e-4 + |x|)
            # with the norm over all dims but the batch one, i.e. (frames, D) for per-frame modulation
            b_, f_ = plan.mod_groups
            m = ws["mod"][:, :b_ * f_].view(-1, b_, f_, 6, D)
            for j in (0, 1, 3, 4):
                x = m[:, :, :, j, :]
                x.div_(torch.linalg.vector_norm(x, dim=(2, 3), keepdim=True, dtype=torch.float32) + 1e-4)
        tail = _Plan()
        tail.steps = rest
        tail.run(stream, self.profile)

    # ---- teacher sampling ---------------------------------------------------------------
    @property
    def unipc_scheduler(self):
        """The multistep scheduler of ``sample()``.  The reference builds diffusers' ``UniPCMultistepScheduler`` from the
        model's scheduler folder (Wan/network.py:635-639); here it defaults to the restated flow-matching UniPC of
        ``fastgen_b200/unipc.py`` and can be assigned any object with the same interface (e.g. the diffusers one)."""
        if getattr(self, "_unipc_scheduler", None) is None:
            from .unipc import FlowUniPCMultistepScheduler

            self._unipc_scheduler = FlowUniPCMultistepScheduler()
        return self._unipc_scheduler

    @unipc_scheduler.setter
    def unipc_scheduler(self, scheduler) -> None:
        self._unipc_scheduler = scheduler

    def _check_fwd_args(self, feature_indices, return_features_early, return_logvar, fwd_pred_type, r):
        if feature_indices is None:
            feature_indices = {}
        if return_features_early and len(feature_indices) == 0:
            return None, True
        if fwd_pred_type is None:
            fwd_pred_type = self.net_pred_type
        else:
            assert fwd_pred_type in NET_PRED_TYPES, f"{fwd_pred_type} is not supported as fwd_pred_type"
        if return_features_early or len(feature_indices) or return_logvar:
            raise NotImplementedError("feature / logvar outputs are training-time paths of the reference")
        if r is not None and self.transformer.r_embedder is None:
            raise ValueError("r_timestep provided but no r_embedder is present")
        return fwd_pred_type, False


class Wan(_WanBase, FastGenNetwork):
    """B200 drop-in for ``fastgen.networks.Wan.network.Wan`` (bidirectional, global AdaLN)."""

    MODEL_ID_14B = "Wan-AI/Wan2.1-T2V-14B-Diffusers"
    MODEL_ID_1_3B = "Wan-AI/Wan2.1-T2V-1.3B-Diffusers"

    def __init__(self, model_id_or_local_path: str = MODEL_ID_1_3B, r_timestep: bool = False,
                 disable_efficient_attn: bool = False, disable_grad_ckpt: bool = False,
                 enable_logvar_linear: bool = True, r_embedder_init: str = "pretrained", time_cond_type: str = "diff",
                 norm_temb: bool = False, net_pred_type: str = "flow", schedule_type: str = "rf",
                 encoder_depth: Optional[int] = None, load_pretrained: bool = False, use_fsdp_checkpoint: bool = True,
                 use_wan_official_sinusoidal: bool = False, transformer_config: Optional[Dict[str, Any]] = None,
                 **model_kwargs):
        FastGenNetwork.__init__(self, net_pred_type=net_pred_type, schedule_type=schedule_type, **model_kwargs)
        if load_pretrained:
            raise NotImplementedError("no hub access: build with load_pretrained=False and call load_state_dict")
        self._init_wan(model_id_or_local_path, transformer_config, enable_logvar_linear, use_wan_official_sinusoidal,
                       r_timestep, norm_temb, encoder_depth, r_embedder_init, time_cond_type)

    def forward(self, x_t: torch.Tensor, t: torch.Tensor, condition: Any = None, r: Optional[torch.Tensor] = None,
                return_features_early: bool = False, feature_indices: Optional[Set[int]] = None,
                return_logvar: bool = False, fwd_pred_type: Optional[str] = None, skip_layers=None,
                unpatchify_features: bool = True, **fwd_kwargs):
        fwd_pred_type, early = self._check_fwd_args(feature_indices, return_features_early, return_logvar, fwd_pred_type, r)
        if early:
            return []
        if skip_layers:
            raise NotImplementedError("skip-layer guidance is a teacher-sampling feature")
        with torch.no_grad():
            condition = torch.stack(condition, dim=0) if isinstance(condition, list) else condition
            t = torch.atleast_1d(t).to(x_t.device)
            if t.numel() == 1 and x_t.shape[0] > 1:
                t = t.expand(x_t.shape[0])
            timestep = self.noise_scheduler.rescale_t(t)
            r_timestep = None
            if r is not None:
                r = torch.atleast_1d(r).to(x_t.device)
                r_timestep = self.noise_scheduler.rescale_t(r.expand(x_t.shape[0]) if r.numel() == 1 else r)
            out = self._run_transformer(x_t.contiguous(), timestep, condition, per_frame=False,
                                        cache_tag=getattr(self, "_bidir_cache_tag", "__bidir__"), r_timestep=r_timestep)
            if fwd_pred_type == self.net_pred_type:
                out = out.clone()   # the caller owns what it gets back (reference: a fresh tensor); `out` is the plan's buffer
            return self.noise_scheduler.convert_model_output(x_t, out, t, src_pred_type=self.net_pred_type,
                                                             target_pred_type=fwd_pred_type)


    def sample(self, noise: torch.Tensor, condition: Any = None, neg_condition: Any = None,
               guidance_scale: Optional[float] = 5.0, num_steps: int = 50, shift: float = 5.0,
               skip_layers: Optional[List[int]] = None, skip_layers_start_percent: float = 0.0, **kwargs) -> torch.Tensor:
        """Teacher multistep sampling with classifier-free guidance: the loop of reference ``Wan.sample``
        (Wan/network.py:919-988) around ``self.unipc_scheduler``.  Two forwards per step (condition, negative
        condition), ``flow = uncond + g (cond - uncond)``, one scheduler step.  ``skip_layers`` (skip-layer guidance) is
        handed to the forward like the reference does, and is not implemented there."""
        assert self.schedule_type == "rf", f"{self.schedule_type} is not supported"
        sch = self.unipc_scheduler
        sch.config.flow_shift = shift
        sch.set_timesteps(num_inference_steps=num_steps, device=noise.device)
        timesteps = sch.timesteps
        t_init = timesteps[0] / sch.config.num_train_timesteps
        latents = self.noise_scheduler.latents(noise=noise, t_init=t_init)
        ns = self.noise_scheduler
        # every t of the trajectory checked once on the host (one copy), not per forward on the device
        t_all = ns.safe_clamp(timesteps / sch.config.num_train_timesteps, min=ns.min_t, max=ns.max_t).to(latents.dtype)
        with ns.validated_t(t_all.cpu()):
            for idx, timestep in enumerate(timesteps):
                t = (timestep / sch.config.num_train_timesteps).expand(latents.shape[0])
                t = ns.safe_clamp(t, min=ns.min_t, max=ns.max_t).to(latents.dtype)
                flow_pred = self(latents, t, condition=condition, r=None, return_features_early=False, feature_indices={},
                                 return_logvar=False)
                if guidance_scale is not None:
                    # its own cache tag: the text keys / values of the negative prompt are computed once, not per step
                    # (an attribute, not a keyword: the forward is called with exactly the reference's arguments)
                    self._bidir_cache_tag = "__bidir_neg__"
                    try:
                        flow_uncond = self(latents, t, condition=neg_condition, r=None, return_features_early=False,
                                           feature_indices={}, return_logvar=False,
                                           skip_layers=skip_layers if idx >= skip_layers_start_percent * num_steps else None)
                    finally:
                        self._bidir_cache_tag = "__bidir__"
                    flow_pred = flow_uncond + guidance_scale * (flow_pred - flow_uncond)
                latents = sch.step(flow_pred, timestep, latents, return_dict=False)[0]
        if self.sp_group is not None and self.sp_exchange == "peer":
            self.check_peer_exchange()
        return latents


class CausalWan(_WanBase, CausalFastGenNetwork):
    """B200 drop-in for ``fastgen.networks.Wan.network_causal.CausalWan`` (per-frame AdaLN, KV-cached chunks)."""

    MODEL_ID_14B = Wan.MODEL_ID_14B
    MODEL_ID_1_3B = Wan.MODEL_ID_1_3B

    def __init__(self, model_id_or_local_path: str = Wan.MODEL_ID_1_3B, r_timestep: bool = False,
                 disable_efficient_attn: bool = False, disable_grad_ckpt: bool = False,
                 enable_logvar_linear: bool = True, r_embedder_init: str = "pretrained", time_cond_type: str = "diff",
                 norm_temb: bool = False, net_pred_type: str = "flow", schedule_type: str = "rf",
                 encoder_depth: Optional[int] = None, load_pretrained: bool = False, use_fsdp_checkpoint: bool = True,
                 chunk_size: int = 3, total_num_frames: int = 21, delete_cache_on_clear: bool = False,
                 use_wan_official_sinusoidal: bool = False, transformer_config: Optional[Dict[str, Any]] = None,
                 **model_kwargs):
        CausalFastGenNetwork.__init__(self, net_pred_type=net_pred_type, schedule_type=schedule_type,
                                      chunk_size=chunk_size, total_num_frames=total_num_frames, **model_kwargs)
        if load_pretrained:
            raise NotImplementedError("no hub access: build with load_pretrained=False and call load_state_dict")
        self._delete_cache_on_clear = delete_cache_on_clear
        self._init_wan(model_id_or_local_path, transformer_config, enable_logvar_linear, use_wan_official_sinusoidal,
                       r_timestep, norm_temb, encoder_depth, r_embedder_init, time_cond_type)

    def clear_caches(self) -> None:
        if self._delete_cache_on_clear:
            self._kv, self._cross, self._plans = {}, {}, {}
            torch.cuda.empty_cache()
            return
        _WanBase.clear_caches(self)

    def forward(self, x_t: torch.Tensor, t: torch.Tensor, condition: Any = None, r: Optional[torch.Tensor] = None,
                return_features_early: bool = False, feature_indices: Optional[Set[int]] = None,
                return_logvar: bool = False, unpatchify_features: bool = True, fwd_pred_type: Optional[str] = None,
                skip_layers=None, cache_tag: str = "pos", cur_start_frame: int = 0, store_kv: bool = False,
                is_ar: bool = False, **fwd_kwargs):
        fwd_pred_type, early = self._check_fwd_args(feature_indices, return_features_early, return_logvar, fwd_pred_type, r)
        if early:
            return []
        if skip_layers:
            raise NotImplementedError("skip-layer guidance is a teacher-sampling feature")
        with torch.no_grad():
            condition = torch.stack(condition, dim=0) if isinstance(condition, list) else condition
            B, _, F, Hh, Ww = x_t.shape
            t = torch.atleast_1d(t).to(x_t.device)
            if t.ndim == 1 and t.numel() == 1 and B > 1:
                t = t.expand(B)
            timestep = self.noise_scheduler.rescale_t(t)
            if timestep.ndim == 1:
                timestep = timestep.view(-1, 1)
            timestep = timestep.expand(B, F) if timestep.shape[1] != F else timestep   # mask of ones * timestep
            use_cache = F < self.total_num_frames
            mask = None
            _, ph, pw = self.cfg["patch"]
            fs = (Hh // ph) * (Ww // pw)
            if not use_cache and not is_ar:
                nc, rem = F // self.chunk_size, F % self.chunk_size
                first = (self.chunk_size + rem) * fs if nc else rem * fs
                mask = (first, self.chunk_size * fs)
            r_timestep = None
            if r is not None:
                r = torch.atleast_1d(r).to(x_t.device)
                if r.ndim == 1 and r.numel() == 1 and B > 1:
                    r = r.expand(B)
                r_timestep = self.noise_scheduler.rescale_t(r)
                r_timestep = r_timestep.view(-1, 1) if r_timestep.ndim == 1 else r_timestep
                r_timestep = r_timestep.expand(B, F) if r_timestep.shape[1] != F else r_timestep
            out = self._run_transformer(x_t.contiguous(), timestep, condition, per_frame=True, cache_tag=cache_tag,
                                        use_cache=use_cache, cur_start_frame=cur_start_frame, store_kv=store_kv,
                                        total_num_frames=self.total_num_frames, mask=mask, r_timestep=r_timestep)
            if fwd_pred_type == self.net_pred_type:
                out = out.clone()   # the caller owns what it gets back (reference: a fresh tensor); `out` is the plan's buffer
            t_conv = t[:, None, :, None, None] if t.ndim == 2 else t
            return self.noise_scheduler.convert_model_output(x_t, out, t_conv, src_pred_type=self.net_pred_type,
                                                             target_pred_type=fwd_pred_type)

    def sample(self, noise: torch.Tensor, condition: Any = None, neg_condition: Any = None,
               guidance_scale: Optional[float] = 5.0, sample_steps: Optional[int] = 50, shift: float = 5.0,
               context_noise: float = 0, **kwargs) -> torch.Tensor:
        """Autoregressive teacher sampling with classifier-free guidance: the loop of reference ``CausalWan.sample``
        (Wan/network_causal.py:1186-1295).  Per chunk (the first one absorbs the remainder frames) a full multistep
        trajectory with two KV-cached forwards per step (cache tags "pos" / "neg"), the chunk written back into ``noise``
        in place like the reference, then one cache-fill forward per tag at t = 0 (or ``context_noise``)."""
        assert self.schedule_type == "rf", f"{self.schedule_type} is not supported"
        was_training = self.training
        self.eval()
        self.clear_caches()
        x = noise
        batch_size, num_frames = x.shape[0], x.shape[2]
        num_chunks, remaining_size = num_frames // self.chunk_size, num_frames % self.chunk_size
        sch = self.unipc_scheduler
        time_rescale_factor = sch.config.num_train_timesteps
        sch.config.flow_shift = shift
        sch.set_timesteps(num_inference_steps=sample_steps, device=noise.device)
        seen = [(sch.timesteps / time_rescale_factor).cpu(), torch.tensor([context_noise if context_noise > 0 else 0.0],
                                                                          dtype=x.dtype)]
        with self.noise_scheduler.validated_t(*seen):   # one host-side check instead of a sync per forward
            for i in range(max(1, num_chunks)):
                if num_chunks == 0:
                    start, end = 0, remaining_size
                else:
                    start = 0 if i == 0 else self.chunk_size * i + remaining_size
                    end = self.chunk_size * (i + 1) + remaining_size
                x_next = x[:, :, start:end]
                sch.config.flow_shift = shift
                sch.set_timesteps(num_inference_steps=sample_steps, device=noise.device)
                for timestep in sch.timesteps:
                    t = (timestep / time_rescale_factor).expand(batch_size)
                    x_cur = x_next
                    flow_pred = self(x_cur, t, condition=condition, cache_tag="pos", cur_start_frame=start, store_kv=False,
                                     is_ar=True)
                    if guidance_scale is not None:
                        flow_uncond = self(x_cur, t, condition=neg_condition, cache_tag="neg", cur_start_frame=start,
                                           store_kv=False, is_ar=True)
                        flow_pred = flow_uncond + guidance_scale * (flow_pred - flow_uncond)
                    x_next = sch.step(flow_pred, timestep, x_next, return_dict=False)[0]
                x[:, :, start:end] = x_next
                # rerun at t = 0 (or with context noise) to write this chunk's keys / values into the caches
                x_cache = x_next
                t_cache = torch.full((batch_size,), 0, device=x.device, dtype=x.dtype)
                if context_noise > 0:
                    t_cache = torch.full((batch_size,), context_noise, device=x.device, dtype=x.dtype)
                    x_cache = self.noise_scheduler.forward_process(x_next, torch.randn_like(x_next), t_cache)
                _ = self(x_cache, t_cache, condition=condition, cache_tag="pos", cur_start_frame=start, store_kv=True, is_ar=True)
                if guidance_scale is not None:
                    _ = self(x_cache, t_cache, condition=neg_condition, cache_tag="neg", cur_start_frame=start, store_kv=True,
                             is_ar=True)
        self.clear_caches()
        self.train(was_training)
        return x
