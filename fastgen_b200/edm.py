"""EDM networks (SongUNet / DhariwalUNet behind EDM preconditioning) on the B200 kernel library.

Drop-in for ``fastgen.networks.EDM.network.EDMPrecond`` (reference EDM/network.py:808-974): same constructor
keywords, same ``forward`` signature, same ``state_dict`` key names (``model.enc.64x64_block0.conv0.weight`` ...),
so checkpoints and the reference's ``generator_fn`` / inference scripts work unchanged.  Inference only.

How a forward is executed (nothing below runs in PyTorch except allocating buffers and a few [B]-sized scalars):
  * the module walks its block table ONCE per batch size and builds a *plan*: pre-allocated channels-last buffers
    and a flat list of C-ABI launches with pre-filled parameter structs;
  * activations between blocks (the residual stream) are fp32 channels-last; every tensor-core operand is a bf16
    channels-last tensor written by the GroupNorm kernel; decoder concatenations are zero-copy (the producer
    writes straight into its slice of the consumer's buffer, kernels take a pixel stride);
  * one UNetBlock = gn_stats, gn_apply(+SiLU,+resample,+raw copy), conv_gemm(+bias,+emb), gn_stats,
    gn_apply(+scale/shift,+SiLU), conv_gemm(conv1 [+ 1x1 skip as a second K segment] + residual, *skip_scale)
    and, with attention, gn_stats, gn_apply, conv_gemm(qkv), attention, conv_gemm(proj + residual, *skip_scale);
  * ``compute="bf16"`` is the throughput path (rel-L2 <= 1e-2 vs the fp32 reference); ``compute="bf16x3"``
    splits every operand into bf16 hi+lo parts and accumulates hi*hi + lo*hi + hi*lo in the same TMEM tile
    (~16 mantissa bits): the fp32 parity path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import ops
from .network import FastGenNetwork
from .noise_schedule import NET_PRED_TYPES

__all__ = ["EDMPrecond", "EDMPrecondB200"]


# ---------------------------------------------------------------------------------------------------------
# parameter containers with the reference's leaf names
# ---------------------------------------------------------------------------------------------------------
def _weight_init(shape: Sequence[int], mode: str, fan_in: int, fan_out: int) -> torch.Tensor:
    """Initialiser families of the reference (EDM/network.py:22-31)."""
    if mode == "xavier_uniform":
        return math.sqrt(6 / (fan_in + fan_out)) * (torch.rand(*shape) * 2 - 1)
    if mode == "xavier_normal":
        return math.sqrt(2 / (fan_in + fan_out)) * torch.randn(*shape)
    if mode == "kaiming_uniform":
        return math.sqrt(3 / fan_in) * (torch.rand(*shape) * 2 - 1)
    if mode == "kaiming_normal":
        return math.sqrt(1 / fan_in) * torch.randn(*shape)
    raise ValueError(f'Invalid init mode "{mode}"')


class _Linear(nn.Module):
    def __init__(self, cin: int, cout: int, bias: bool = True, init_mode: str = "kaiming_normal",
                 init_weight: float = 1, init_bias: float = 0):
        super().__init__()
        self.in_features, self.out_features = cin, cout
        kw = dict(mode=init_mode, fan_in=cin, fan_out=cout)
        self.weight = nn.Parameter(_weight_init([cout, cin], **kw) * init_weight)
        self.bias = nn.Parameter(_weight_init([cout], **kw) * init_bias) if bias else None


class _Conv(nn.Module):
    def __init__(self, cin: int, cout: int, kernel: int, bias: bool = True, up: bool = False, down: bool = False,
                 resample_filter: Sequence[int] = (1, 1), init_mode: str = "kaiming_normal", init_weight: float = 1,
                 init_bias: float = 0):
        super().__init__()
        assert not (up and down)
        self.in_channels, self.out_channels, self.kernel, self.up, self.down = cin, cout, kernel, up, down
        kw = dict(mode=init_mode, fan_in=cin * kernel * kernel, fan_out=cout * kernel * kernel)
        self.weight = nn.Parameter(_weight_init([cout, cin, kernel, kernel], **kw) * init_weight) if kernel else None
        self.bias = nn.Parameter(_weight_init([cout], **kw) * init_bias) if kernel and bias else None
        f = torch.as_tensor(list(resample_filter), dtype=torch.float32)
        f = f.ger(f).unsqueeze(0).unsqueeze(1) / f.sum().square()
        # kept only so that state_dict() has the reference's keys; the [1,1] box filter is baked into the kernels
        self.register_buffer("resample_filter", f if up or down else None)


class _GroupNorm(nn.Module):
    def __init__(self, c: int, num_groups: int = 32, min_channels_per_group: int = 4, eps: float = 1e-5):
        super().__init__()
        self.num_groups = min(num_groups, c // min_channels_per_group)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _Block(nn.Module):
    """Parameters of one UNetBlock (reference EDM/network.py:205-272)."""

    def __init__(self, cin: int, cout: int, emb: int, up=False, down=False, attention=False, num_heads=None,
                 channels_per_head=64, dropout=0.0, skip_scale=1.0, eps=1e-5, resample_filter=(1, 1),
                 resample_proj=False, adaptive_scale=True, init=None, init_zero=None, init_attn=None):
        super().__init__()
        init = init or {}
        init_zero = init_zero if init_zero is not None else dict(init_weight=0)
        self.in_channels, self.out_channels, self.emb_channels = cin, cout, emb
        self.num_heads = 0 if not attention else num_heads if num_heads is not None else cout // channels_per_head
        self.dropout, self.skip_scale, self.adaptive_scale = dropout, skip_scale, adaptive_scale
        self.up, self.down = up, down
        self.norm0 = _GroupNorm(cin, eps=eps)
        self.conv0 = _Conv(cin, cout, 3, up=up, down=down, resample_filter=resample_filter, **init)
        self.affine = _Linear(emb, cout * (2 if adaptive_scale else 1), **init)
        self.norm1 = _GroupNorm(cout, eps=eps)
        self.conv1 = _Conv(cout, cout, 3, **init_zero)
        self.skip = None
        if cout != cin or up or down:
            kernel = 1 if resample_proj or cout != cin else 0
            self.skip = _Conv(cin, cout, kernel, up=up, down=down, resample_filter=resample_filter, **init)
        if self.num_heads:
            self.norm2 = _GroupNorm(cout, eps=eps)
            self.qkv = _Conv(cout, cout * 3, 1, **(init_attn if init_attn is not None else init))
            self.proj = _Conv(cout, cout, 1, **init_zero)


class _UNet(nn.Module):
    """Parameter tree + block table of SongUNet / DhariwalUNet (reference EDM/network.py:346-487, 584-678)."""

    def __init__(self, model_type: str, img_resolution: int, in_channels: int, out_channels: int, label_dim=0,
                 augment_dim=0, model_channels=None, channel_mult=None, channel_mult_emb=4, num_blocks=None,
                 attn_resolutions=None, dropout=0.10, label_dropout=0, embedding_type="positional",
                 channel_mult_noise=1, encoder_type="standard", decoder_type="standard", resample_filter=(1, 1),
                 r_timestep=False, **kwargs):
        super().__init__()
        self.model_type = model_type
        song = model_type == "SongUNet"
        if model_type not in ("SongUNet", "DhariwalUNet"):
            raise KeyError(model_type)
        model_channels = model_channels if model_channels is not None else (128 if song else 192)
        channel_mult = list(channel_mult) if channel_mult is not None else ([1, 2, 2, 2] if song else [1, 2, 3, 4])
        num_blocks = num_blocks if num_blocks is not None else (4 if song else 3)
        attn_resolutions = list(attn_resolutions) if attn_resolutions is not None else ([16] if song else [32, 16, 8])
        if song and (embedding_type != "positional" or encoder_type != "standard" or decoder_type != "standard"):
            raise NotImplementedError("only the DDPM++ flavour (positional / standard / standard) is built")
        if list(resample_filter) != [1, 1]:
            raise NotImplementedError("only the [1,1] resample filter is built")
        self.img_resolution, self.in_channels, self.out_channels = img_resolution, in_channels, out_channels
        self.label_dim, self.augment_dim, self.label_dropout = label_dim, augment_dim, label_dropout
        self.has_r = bool(r_timestep)
        emb_ch = model_channels * channel_mult_emb
        self.emb_channels = emb_ch
        if song:
            self.noise_channels = model_channels * channel_mult_noise
            cond_ch = self.noise_channels * (1 + self.has_r)
            init = dict(init_mode="xavier_uniform")
            init_zero = dict(init_mode="xavier_uniform", init_weight=1e-5)
            init_attn = dict(init_mode="xavier_uniform", init_weight=math.sqrt(0.2))
            bk = dict(emb=emb_ch, num_heads=1, dropout=dropout, skip_scale=math.sqrt(0.5), eps=1e-6,
                      resample_filter=resample_filter, resample_proj=True, adaptive_scale=False, init=init,
                      init_zero=init_zero, init_attn=init_attn)
            self.map_label = _Linear(label_dim, cond_ch, **init) if label_dim else None
            self.map_augment = _Linear(augment_dim, cond_ch, bias=False, **init) if augment_dim else None
            self.map_layer0 = _Linear(cond_ch, emb_ch, **init)
            self.map_layer1 = _Linear(emb_ch, emb_ch, **init)
            self.final_eps = 1e-6
        else:
            self.noise_channels = model_channels
            cond_ch = model_channels * (1 + self.has_r)
            init = dict(init_mode="kaiming_uniform", init_weight=math.sqrt(1 / 3), init_bias=math.sqrt(1 / 3))
            init_zero = dict(init_mode="kaiming_uniform", init_weight=0, init_bias=0)
            bk = dict(emb=emb_ch, channels_per_head=64, dropout=dropout, init=init, init_zero=init_zero)
            self.map_augment = _Linear(augment_dim, cond_ch, bias=False, **init_zero) if augment_dim else None
            self.map_layer0 = _Linear(cond_ch, emb_ch, **init)
            self.map_layer1 = _Linear(emb_ch, emb_ch, **init)
            self.map_label = (_Linear(label_dim, emb_ch, bias=False, init_mode="kaiming_normal",
                                      init_weight=math.sqrt(label_dim)) if label_dim else None)
            self.final_eps = 1e-5

        # ---- encoder ----
        self.enc = nn.ModuleDict()
        cout = in_channels
        for level, mult in enumerate(channel_mult):
            res = img_resolution >> level
            if level == 0:
                cin, cout = cout, (model_channels if song else model_channels * mult)
                self.enc[f"{res}x{res}_conv"] = _Conv(cin, cout, 3, **init)
            else:
                self.enc[f"{res}x{res}_down"] = _Block(cout, cout, down=True, **bk)
            for idx in range(num_blocks):
                cin, cout = cout, model_channels * mult
                self.enc[f"{res}x{res}_block{idx}"] = _Block(cin, cout, attention=(res in attn_resolutions), **bk)
        skips = [m.out_channels for m in self.enc.values()]

        # ---- decoder ----
        self.dec = nn.ModuleDict()
        self.concat_with_skip: Dict[str, int] = {}  # decoder block name -> channels of the skip it concatenates
        for level, mult in reversed(list(enumerate(channel_mult))):
            res = img_resolution >> level
            if level == len(channel_mult) - 1:
                self.dec[f"{res}x{res}_in0"] = _Block(cout, cout, attention=True, **bk)
                self.dec[f"{res}x{res}_in1"] = _Block(cout, cout, **bk)
            else:
                self.dec[f"{res}x{res}_up"] = _Block(cout, cout, up=True, **bk)
            for idx in range(num_blocks + 1):
                cs = skips.pop()
                cin, cout = cout + cs, model_channels * mult
                attn = (idx == num_blocks and res in attn_resolutions) if song else (res in attn_resolutions)
                name = f"{res}x{res}_block{idx}"
                self.dec[name] = _Block(cin, cout, attention=attn, **bk)
                self.concat_with_skip[name] = cs
            if song and level == 0:
                self.dec[f"{res}x{res}_aux_norm"] = _GroupNorm(cout, eps=1e-6)
                self.dec[f"{res}x{res}_aux_conv"] = _Conv(cout, out_channels, 3, **init_zero)
        if not song:
            self.out_norm = _GroupNorm(cout)
            self.out_conv = _Conv(cout, out_channels, 3, **init_zero)
        self.logvar_linear = _Linear(self.noise_channels, 1)

    # final norm / conv under either naming
    def final_layers(self) -> Tuple[_GroupNorm, _Conv]:
        if self.model_type == "SongUNet":
            r = self.img_resolution
            return self.dec[f"{r}x{r}_aux_norm"], self.dec[f"{r}x{r}_aux_conv"]
        return self.out_norm, self.out_conv


# ---------------------------------------------------------------------------------------------------------
# execution plan
# ---------------------------------------------------------------------------------------------------------
# tests set this to build launch lists against host memory so the argument validation of every C-ABI call (and the
# algorithmic byte / FLOP bookkeeping) can be exercised on a CPU-only host; such a plan cannot execute.
_ALLOW_HOST_PLANS = False


@dataclass
class _Act:
    """channels-last activation view: channels [off, off+c) of a buffer [n, h, w, ld]."""
    buf: torch.Tensor
    off: int
    c: int

    @property
    def n(self):
        return self.buf.shape[0]

    @property
    def h(self):
        return self.buf.shape[1]

    @property
    def w(self):
        return self.buf.shape[2]

    @property
    def ld(self):
        return self.buf.shape[3]

    @property
    def t(self) -> torch.Tensor:
        return self.buf[..., self.off:self.off + self.c]


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def _split(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


_X3_ATTENTION_DEFAULT = os.environ.get("FASTGEN_B200_X3_ATTENTION", "f32")
_F32_ATTENTION_WIDTHS = (64, 128, 192, 256)   # head widths fg_attention_f32 is instantiated for


class _Plan:
    """Buffers + the flat launch list for one (batch size, dtype of x_t is irrelevant) forward."""

    def __init__(self, net: "EDMPrecond", n: int):
        self.net, self.n = net, n
        self.dev = next(net.parameters()).device
        self.x3 = net.compute == "bf16x3"
        self.h1_bf16 = (not self.x3) and bool(net.bf16_h1)
        self.attn_f32 = self.x3 and net.x3_attention == "f32"   # fp32 CUDA-core attention on the fp32-parity path
        self._warned_widths: Set[int] = set()
        self.steps: List[Tuple[Any, tuple]] = []
        self.step_flops: List[float] = []
        self.step_desc: List[Tuple[str, float]] = []
        self.keep: List[Any] = []
        self.runs = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._build()

    # ---- buffer helpers ----
    def _new(self, shape, dtype) -> torch.Tensor:
        t = torch.zeros(shape, device=self.dev, dtype=dtype)
        self.keep.append(t)
        return t

    def _size_scratch(self) -> None:
        """One flat buffer per scratch role, sized for the largest block; blocks run in order on one stream, so
        they can all share them."""
        m, n = self.net.model, self.n
        need: Dict[str, int] = {}

        def req(name, numel):
            need[name] = max(need.get(name, 0), numel)

        for md in (m.enc, m.dec):
            for name, b in md.items():
                if not isinstance(b, _Block):
                    continue
                r = int(name.split("x")[0])
                px = n * r * r
                req("h0", px * b.in_channels)
                if b.skip is not None and b.skip.kernel == 1:
                    req("raw", px * b.in_channels)
                elif b.skip is not None and b.down:
                    req("xds", px * b.in_channels)
                req("h1", px * b.out_channels)
                req("h2", px * b.out_channels)
                if b.num_heads:
                    req("mid", px * b.out_channels)
                    req("h3", px * b.out_channels)
                    req("qkv_f32" if self._f32_attention(b) else "qkv", px * 3 * b.out_channels)
                    req("att", px * b.out_channels)
                    if self._f32_attention(b):
                        req("att_lo", px * b.out_channels)
        R = m.img_resolution
        last = list(m.dec.values())
        c_last = [b for b in last if isinstance(b, _Block)][-1].out_channels
        req("h0", n * R * R * c_last)
        # h1 (conv0's output) is read by norm1 only: on the bf16 path it is stored as bf16 (statistics of the rounded values
        # from the producer's epilogue), which takes a third off norm1's HBM traffic and half off conv0's stores
        f32_roles = ("mid", "qkv_f32", "xds") if self.h1_bf16 else ("h1", "mid", "qkv_f32", "xds")
        self._scratch_buf: Dict[str, torch.Tensor] = {}
        for name, numel in need.items():
            dt = torch.float32 if name in f32_roles else torch.bfloat16
            self._scratch_buf[name] = self._new([numel], dt)
            if self.x3 and name in ("h0", "raw", "h2", "h3"):
                self._scratch_buf[name + "_lo"] = self._new([numel], dt)

    def _f32_attention(self, b: "_Block") -> bool:
        """fp32 CUDA-core attention for this block: the split-precision path, unless switched off or the head width has
        no fg_attention_f32 instance (then the bf16 tensor-core kernel runs, ~1e-3 instead of ~1e-5)"""
        if not (self.attn_f32 and b.num_heads):
            return False
        d = b.out_channels // b.num_heads
        if d not in _F32_ATTENTION_WIDTHS:
            if d not in self._warned_widths:
                self._warned_widths.add(d)
                warnings.warn(f"compute='bf16x3': no fp32 attention kernel for head width {d} (instances: "
                              f"{_F32_ATTENTION_WIDTHS}); this block uses the bf16 tensor-core attention (~1e-3 instead of "
                              "~1e-5 of the fp32 reference)")
            return False
        return True

    def _scratch(self, name: str, shape) -> torch.Tensor:
        numel = int(np.prod(shape))
        return self._scratch_buf[name][:numel].view(*shape)

    # ---- launch helpers (append to the list; tensors referenced by structs are kept alive) ----
    def _launch(self, fn, *args, flops: float = 0.0, desc: str = "", nbytes: float = 0.0):
        self.steps.append((fn, args))
        self.step_flops.append(flops)
        self.step_desc.append((desc, nbytes))

    def conv(self, segs, n, h, w, cout, out: torch.Tensor, out_ld: int, **kw):
        p = ops.conv_gemm_params(segs, n, h, w, cout, out, out_ld, **kw)
        self.keep.append((p, segs, out, kw))
        flops = 2.0 * n * h * w * cout * sum(s.cin * s.taps for s in segs if s.alg)
        ktot = sum(s.cin * s.taps for s in segs if s.alg)
        res = kw.get("res")
        nbytes = n * h * w * (sum(2 * s.cin for s in segs if s.alg) + cout * out.element_size()
                              + (cout * res.element_size() if res is not None else 0))
        fn, args, keep = ops.conv_gemm_launch(p)
        self.keep.append(keep)
        self._launch(fn, *args, flops=flops, nbytes=nbytes,
                     desc=f"conv {h}x{w} K={ktot} N={cout} out={str(out.dtype)[6:]} res={kw.get('res_mode', 0)}")

    def new_stats(self) -> torch.Tensor:
        """one [n, 32, 2] fp64 slot of the statistics arena (zeroed by the first step of the plan)"""
        i = self._stats_used
        self._stats_used += 1
        assert i < self.stats_arena.shape[0], "planner bug: statistics arena too small"
        return self.stats_arena[i]

    def fused_stats(self, h: int, w: int, cout: int, out_ld: int, norm: "_GroupNorm", c_norm: int, bf16_out: bool = False):
        """kwargs that make a conv accumulate the GroupNorm statistics `norm` needs of its output, or None when the
        producer cannot (tiny images, narrow or unaligned outputs): the consumer then runs fg_gn_stats itself."""
        if not self.net.fuse_gn_stats or h * w < 32 or cout % 4 or cout <= 16 or out_ld % 4 or cout != c_norm:
            return None
        if bf16_out and ((c_norm // norm.num_groups) % 2 or out_ld % 8):   # the epilogue sums column pairs of a bf16 slab
            return None
        slot = self.new_stats()
        return dict(stats=slot, stats_groups=norm.num_groups, stats_cg=c_norm // norm.num_groups, stats_ch0=0)

    def gn(self, x: _Act, norm: _GroupNorm, pk: Dict[str, torch.Tensor], y: torch.Tensor, silu: bool, resample: int = 0,
           scale=None, shift=None, mod_ld=0, y_lo=None, raw=None, raw_lo=None, stats: Optional[torch.Tensor] = None,
           raw_f32=None):
        n, h, w, c = x.n, x.h, x.w, x.c
        groups = norm.num_groups
        xv = x.t
        if stats is None:   # not produced by the conv that wrote x: one streaming pass over x
            stats = self.new_stats()
            self._launch(L.lib().fg_gn_stats, xv.data_ptr(), ops._DT[xv.dtype], x.ld, n, h * w, c, groups,
                         stats.data_ptr(), desc=f"gn_stats {h}x{w} c={c}", nbytes=n * h * w * c * xv.element_size())
        p = ops.gn_apply_params(xv, x.ld, n, h, w, c, groups, stats, norm.eps, pk["gamma"], pk["beta"], y, c,
                                scale=scale, shift=shift, mod_ld=mod_ld, silu=silu, resample=resample, y_lo=y_lo,
                                raw=raw, raw_ld=c, raw_lo=raw_lo, raw_f32=raw_f32, raw_f32_ld=c)
        self.keep.append((p, xv, y, y_lo, raw, raw_lo, scale, shift, stats, raw_f32))
        outs = 1 + (y_lo is not None) + (raw is not None) + (raw_lo is not None)
        opix = h * w * (4 if resample == L.FG_RES_UP2 else 1) // (4 if resample == L.FG_RES_DOWN2 else 1)
        self._launch(L.lib().fg_gn_apply, C.byref(p), desc=f"gn_apply {h}x{w} c={c} outs={outs} rs={resample}",
                     nbytes=n * c * (h * w * xv.element_size() + opix * 2 * outs))

    def segs3(self, a_hi, a_lo, ld, w_hi, w_lo, cin, taps) -> List[ops.Seg]:
        s = [ops.Seg(a_hi, ld, w_hi, cin, taps)]
        if self.x3:
            s += [ops.Seg(a_lo, ld, w_hi, cin, taps, alg=False), ops.Seg(a_hi, ld, w_lo, cin, taps, alg=False)]
        return s

    def linear(self, a_hi, a_lo, k: int, wpk, cout: int, out: torch.Tensor, out_ld: int, **kw):
        """small fp32-accurate GEMM of the embedding path: always split precision (3 segments)."""
        segs = [ops.Seg(a_hi, k, wpk["hi"], k, 1), ops.Seg(a_lo, k, wpk["hi"], k, 1, alg=False),
                ops.Seg(a_hi, k, wpk["lo"], k, 1, alg=False)]
        self.conv(segs, self.n, 1, 1, cout, out, out_ld, bias=wpk.get("bias"), **kw)

    def split(self, x: torch.Tensor, x_ld: int, c: int, act: int = 0):
        hi = self._new([self.n, c], torch.bfloat16)
        lo = self._new([self.n, c], torch.bfloat16)
        self._launch(L.lib().fg_split_bf16, x.data_ptr(), x_ld, None, 0, self.n, c, act, hi.data_ptr(), lo.data_ptr(), c,
                     None, 0)
        return hi, lo

    # ---- the plan ----
    def _build(self):
        net, m, n = self.net, self.net.model, self.n
        pk = net._packed
        song = m.model_type == "SongUNet"
        R = m.img_resolution
        f32, bf16 = torch.float32, torch.bfloat16
        n_blocks = sum(isinstance(b, _Block) for md in (m.enc, m.dec) for b in md.values())
        self.stats_arena = self._new([3 * n_blocks + 4, n, 32, 2], torch.float64)
        self._stats_used = 0
        self._launch(L.lib().fg_zero, self.stats_arena.data_ptr(), self.stats_arena.numel() * 8, desc="zero_stats")
        self._size_scratch()

        # ---------------- inputs staged by forward() ----------------
        self.t_noise = self._new([n], f32)
        self.r_noise = self._new([n], f32) if m.has_r else None
        self.labels = self._new([n, _pad8(m.label_dim)], f32) if m.label_dim else None
        self.augment = self._new([n, _pad8(m.augment_dim)], f32) if m.augment_dim else None
        self.use_augment = False
        cin_pad = _pad8(m.in_channels)
        self.xin = self._new([n, R, R, cin_pad], bf16)
        self.xin_lo = self._new([n, R, R, cin_pad], bf16) if self.x3 else None

        # ---------------- embedding path (fp32-accurate, split precision) ----------------
        nc = m.noise_channels
        cond = nc * (1 + m.has_r)
        e0 = self._new([n, cond], f32)
        lib = L.lib()
        self._launch(lib.fg_pos_embedding, self.t_noise.data_ptr(), n, nc, 10000, int(song), int(song), e0.data_ptr(), cond)
        if m.has_r:
            self._launch(lib.fg_pos_embedding, self.r_noise.data_ptr(), n, nc, 10000, int(song), int(song),
                         e0[:, nc:].data_ptr(), cond)
        emb_ch = m.emb_channels
        cur = e0
        self._augment_steps_at = None
        if song:
            if m.map_label is not None:
                lh, ll = self.split(self.labels, self.labels.shape[1], self.labels.shape[1])
                e0b = self._new([n, cond], f32)
                self.linear(lh, ll, self.labels.shape[1], pk["map_label"], cond, e0b, cond, res=cur, res_ld=cond,
                            res_mode=L.FG_RES_SAME)
                cur = e0b
            if m.map_augment is not None:
                self.aug_hi, self.aug_lo = self.split(self.augment, self.augment.shape[1], self.augment.shape[1])
                e0c = self._new([n, cond], f32)
                self.linear(self.aug_hi, self.aug_lo, self.augment.shape[1], pk["map_augment"], cond, e0c, cond, res=cur,
                            res_ld=cond, res_mode=L.FG_RES_SAME)
                cur = e0c
            h, l = self.split(cur, cond, cond)
            e1 = self._new([n, emb_ch], f32)
            self.linear(h, l, cond, pk["map_layer0"], emb_ch, e1, emb_ch)
            h, l = self.split(e1, emb_ch, emb_ch, act=1)
            e2 = self._new([n, emb_ch], f32)
            self.linear(h, l, emb_ch, pk["map_layer1"], emb_ch, e2, emb_ch)
            emb_hi, emb_lo = self.split(e2, emb_ch, emb_ch, act=1)
        else:
            if m.map_augment is not None:
                self.aug_hi, self.aug_lo = self.split(self.augment, self.augment.shape[1], self.augment.shape[1])
                e0c = self._new([n, cond], f32)
                self.linear(self.aug_hi, self.aug_lo, self.augment.shape[1], pk["map_augment"], cond, e0c, cond, res=cur,
                            res_ld=cond, res_mode=L.FG_RES_SAME)
                cur = e0c
            h, l = self.split(cur, cond, cond)
            e1 = self._new([n, emb_ch], f32)
            self.linear(h, l, cond, pk["map_layer0"], emb_ch, e1, emb_ch)
            h, l = self.split(e1, emb_ch, emb_ch, act=1)
            e2 = self._new([n, emb_ch], f32)
            self.linear(h, l, emb_ch, pk["map_layer1"], emb_ch, e2, emb_ch)
            if m.map_label is not None:
                lh, ll = self.split(self.labels, self.labels.shape[1], self.labels.shape[1])
                e3 = self._new([n, emb_ch], f32)
                self.linear(lh, ll, self.labels.shape[1], pk["map_label"], emb_ch, e3, emb_ch, res=e2, res_ld=emb_ch,
                            res_mode=L.FG_RES_SAME)
                e2 = e3
            emb_hi, emb_lo = self.split(e2, emb_ch, emb_ch, act=1)
        aff_total = pk["affine_all"]["hi"].shape[0]
        self.affine = self._new([n, aff_total], f32)
        self.linear(emb_hi, emb_lo, emb_ch, pk["affine_all"], aff_total, self.affine, aff_total)

        # ---------------- activation homes (zero-copy decoder concatenation) ----------------
        enc_names = list(m.enc.keys())
        dec_names = [k for k in m.dec.keys() if "aux" not in k]
        res_of = lambda name: int(name.split("x")[0])
        # the skip stack: every encoder entry pushes, decoder block{idx} pops
        skip_stack = list(enc_names)
        skip_home: Dict[str, Tuple[torch.Tensor, int]] = {}   # encoder entry -> (concat buffer, channel offset)
        x_home: Dict[str, Tuple[torch.Tensor, int]] = {}      # decoder entry -> where its INPUT x must have been written
        prev_c = m.enc[enc_names[-1]].out_channels
        for name in dec_names:
            blk = m.dec[name]
            r_in = res_of(name) // 2 if blk.up else res_of(name)
            if name in m.concat_with_skip:
                cs = m.concat_with_skip[name]
                cx = blk.in_channels - cs
                buf = self._new([n, r_in, r_in, blk.in_channels], f32)
                x_home[name] = (buf, 0)
                skip_home[skip_stack.pop()] = (buf, cx)
                assert cx == prev_c
            prev_c = blk.out_channels

        def out_act(c: int, res: int, home: Optional[Tuple[torch.Tensor, int]]) -> _Act:
            if home is not None:
                return _Act(home[0], home[1], c)
            return _Act(self._new([n, res, res, c], f32), 0, c)

        # ---------------- encoder / decoder ----------------
        # The conv that writes a tensor also accumulates the GroupNorm statistics of its consumers (fg_conv_gemm stats /
        # stats2): the next entry's norm0 when that norm sees exactly this tensor, and -- for the two producers of a
        # decoder skip concatenation torch.cat([x, skip]) (x at channel 0, the encoder skip at channel cx) -- the
        # concatenation's norm0, each producer adding its channels' share to the same [n, 32, 2] slot.
        entries = [("enc", nm) for nm in enc_names] + [("dec", nm) for nm in dec_names]
        final_norm, final_conv = m.final_layers()

        def can_fuse(px: int, cout: int, ld: int) -> bool:
            return bool(self.net.fuse_gn_stats) and px >= 32 and cout % 4 == 0 and cout > 16 and ld % 4 == 0

        def next_norm(i: int):
            if i + 1 < len(entries):
                side, nm = entries[i + 1]
                if side == "dec" and nm in x_home:
                    return None          # the next block normalises a concatenation
                return m.enc[nm].norm0 if side == "enc" else m.dec[nm].norm0
            return final_norm

        # which decoder block pops which encoder entry, and whether both producers of its input can fuse the statistics
        popped_by: Dict[str, str] = {}
        stack = list(enc_names)
        for name in dec_names:
            if name in m.concat_with_skip:
                popped_by[stack.pop()] = name
        cat_slot: Dict[str, torch.Tensor] = {}
        for ei, (side, name) in enumerate(entries):
            if side != "dec" or name not in x_home:
                continue
            blk = m.dec[name]
            buf, _ = x_home[name]
            px, ld = buf.shape[1] * buf.shape[2], buf.shape[3]
            cs = m.concat_with_skip[name]
            cx = blk.in_channels - cs
            if can_fuse(px, cx, ld) and can_fuse(px, cs, ld) and blk.in_channels % blk.norm0.num_groups == 0:
                cat_slot[name] = self.new_stats()

        x: Optional[_Act] = None
        x_stats: Optional[torch.Tensor] = None
        for i, (side, name) in enumerate(entries):
            mod = m.enc[name] if side == "enc" else m.dec[name]
            res = res_of(name)
            if side == "enc":
                dst = out_act(mod.out_channels, res, skip_home.get(name))
            else:
                if name in x_home:
                    buf, _ = x_home[name]
                    assert x.buf is buf, "planner bug: decoder input was not written into its concat buffer"
                    x = _Act(buf, 0, mod.in_channels)
                    x_stats = cat_slot.get(name)
                di = dec_names.index(name)
                nxt = dec_names[di + 1] if di + 1 < len(dec_names) else None
                dst = out_act(mod.out_channels, res, x_home.get(nxt) if nxt is not None else None)
            # statistics targets of this entry's output
            targets: List[Tuple[torch.Tensor, int, int, int]] = []   # (slot, groups, channels per group, channel offset)
            own_slot = None
            cout = mod.out_channels
            fusable = can_fuse(res * res, cout, dst.ld)
            nn_ = next_norm(i)
            if nn_ is not None and fusable and cout % nn_.num_groups == 0:
                own_slot = self.new_stats()
                targets.append((own_slot, nn_.num_groups, cout // nn_.num_groups, 0))
            nxt_entry = entries[i + 1] if i + 1 < len(entries) else None
            if nxt_entry is not None and nxt_entry[0] == "dec" and nxt_entry[1] in cat_slot:
                cblk = m.dec[nxt_entry[1]]
                targets.append((cat_slot[nxt_entry[1]], cblk.norm0.num_groups, cblk.in_channels // cblk.norm0.num_groups, 0))
            if side == "enc" and popped_by.get(name) in cat_slot:
                cname = popped_by[name]
                cblk = m.dec[cname]
                targets.append((cat_slot[cname], cblk.norm0.num_groups, cblk.in_channels // cblk.norm0.num_groups,
                                cblk.in_channels - m.concat_with_skip[cname]))
            assert len(targets) <= 2, "planner bug: more than two statistics consumers"
            out_stats = None
            if targets:
                t0 = targets[0]
                out_stats = dict(stats=t0[0], stats_groups=t0[1], stats_cg=t0[2], stats_ch0=t0[3])
                if len(targets) == 2:
                    t1 = targets[1]
                    out_stats.update(stats2=t1[0], stats2_groups=t1[1], stats2_cg=t1[2], stats2_ch0=t1[3])
            if isinstance(mod, _Conv):
                w = pk[f"enc.{name}"]
                self.conv(self.segs3(self.xin, self.xin_lo, cin_pad, w["hi"], w.get("lo"), cin_pad, 9), n, R, R,
                          mod.out_channels, dst.t, dst.ld, bias=w["bias"], **(out_stats or {}))
            else:
                self._block(f"{side}.{name}", mod, x, dst, x_stats, out_stats)
            x = dst
            x_stats = own_slot
        # ---------------- output head ----------------
        norm, conv = final_norm, final_conv
        key = "final"
        hb = self._scratch("h0", [n, R, R, x.c])
        hb_lo = self._scratch("h0_lo", [n, R, R, x.c]) if self.x3 else None
        self.gn(x, norm, pk[key + ".norm"], hb, silu=True, y_lo=hb_lo, stats=x_stats)
        self.f_ld = 4 if m.out_channels <= 4 else _pad8(m.out_channels)
        self.f_out = self._new([n, R, R, self.f_ld], f32)
        w = pk[key + ".conv"]
        self.conv(self.segs3(hb, hb_lo, x.c, w["hi"], w.get("lo"), x.c, 9), n, R, R, m.out_channels, self.f_out, self.f_ld,
                  bias=w["bias"])

    def _block(self, key: str, b: _Block, x: _Act, dst: _Act, x_stats: Optional[torch.Tensor] = None,
               out_stats: Optional[Dict[str, Any]] = None):
        """Launch list of one UNetBlock (reference EDM/network.py:274-299).  `x_stats`: norm0 statistics already
        accumulated by the producer of x; `out_stats`: conv kwargs that accumulate the next norm0's statistics."""
        n = self.n
        pk = self.net._packed
        f32, bf16 = torch.float32, torch.bfloat16
        cin, cout = b.in_channels, b.out_channels
        hi_, wi_ = x.h, x.w
        ho, wo = (hi_ * 2, wi_ * 2) if b.up else (hi_ // 2, wi_ // 2) if b.down else (hi_, wi_)
        resample = L.FG_RES_UP2 if b.up else L.FG_RES_DOWN2 if b.down else L.FG_RES_NONE
        skip_conv = b.skip is not None and b.skip.kernel == 1
        # norm0 + SiLU (+ resample) -> conv0 operand; raw (resampled) x -> 1x1 skip operand
        h0 = self._scratch("h0", [n, ho, wo, cin])
        h0_lo = self._scratch("h0_lo", [n, ho, wo, cin]) if self.x3 else None
        raw = self._scratch("raw", [n, ho, wo, cin]) if skip_conv else None
        raw_lo = self._scratch("raw_lo", [n, ho, wo, cin]) if (skip_conv and self.x3) else None
        # a down block without a skip projection adds the 2x2 mean of x: norm0's pass (which reads x anyway) leaves it in fp32,
        # so that conv1 takes it as a same-size residual through its TMA-staged slab ring (plain loads of four pixels per
        # output in the epilogue ran the three such convolutions of the IN-64 network at 200-460 TFLOP/s)
        xds = None
        if b.down and b.skip is not None and not skip_conv:
            xds = self._scratch("xds", [n, ho, wo, cin])
        self.gn(x, b.norm0, pk[key + ".norm0"], h0, silu=True, resample=resample, y_lo=h0_lo, raw=raw, raw_lo=raw_lo,
                stats=x_stats, raw_f32=xds)
        # conv0 (+ bias, + additive embedding)
        aff_off = pk[key + ".affine_off"]
        aff = self.affine[:, aff_off:]
        h1 = self._scratch("h1", [n, ho, wo, cout])
        w0 = pk[key + ".conv0"]
        kw = {}
        if not b.adaptive_scale:
            kw = dict(rowvec=aff, rowvec_ld=self.affine.shape[1])
        st1 = self.fused_stats(ho, wo, cout, cout, b.norm1, cout, bf16_out=self.h1_bf16)
        self.conv(self.segs3(h0, h0_lo, cin, w0["hi"], w0.get("lo"), cin, 9), n, ho, wo, cout, h1, cout, bias=w0["bias"],
                  **kw, **(st1 or {}))
        s1 = st1["stats"] if st1 else None
        # norm1 (+ adaptive scale/shift) + SiLU -> conv1 operand
        h2 = self._scratch("h2", [n, ho, wo, cout])
        h2_lo = self._scratch("h2_lo", [n, ho, wo, cout]) if self.x3 else None
        h1a = _Act(h1, 0, cout)
        if b.adaptive_scale:
            self.gn(h1a, b.norm1, pk[key + ".norm1"], h2, silu=True, scale=aff, shift=self.affine[:, aff_off + cout:],
                    mod_ld=self.affine.shape[1], y_lo=h2_lo, stats=s1)
        else:
            self.gn(h1a, b.norm1, pk[key + ".norm1"], h2, silu=True, y_lo=h2_lo, stats=s1)
        # conv1 (+ 1x1 skip projection as extra K segments | + resampled identity skip), * skip_scale
        w1 = pk[key + ".conv1"]
        segs = self.segs3(h2, h2_lo, cout, w1["hi"], w1.get("lo"), cout, 9)
        bias = w1["bias"]
        rkw = {}
        if skip_conv:
            ws = pk[key + ".skip"]
            segs += self.segs3(raw, raw_lo, cin, ws["hi"], ws.get("lo"), cin, 1)
            bias = pk[key + ".bias1s"]
        else:
            mode = L.FG_RES_SAME if b.skip is None else resample
            rkw = dict(res=x.t, res_ld=x.ld, res_mode=mode)
            if xds is not None:
                rkw = dict(res=xds, res_ld=cin, res_mode=L.FG_RES_SAME)
        mid = dst if not b.num_heads else _Act(self._scratch("mid", [n, ho, wo, cout]), 0, cout)
        st2 = self.fused_stats(ho, wo, cout, mid.ld, b.norm2, cout) if b.num_heads else out_stats
        self.conv(segs, n, ho, wo, cout, mid.t, mid.ld, bias=bias, out_scale=float(b.skip_scale), **rkw, **(st2 or {}))
        if not b.num_heads:
            return
        # attention: norm2 -> qkv (rows permuted to q|k|v, head, d) -> flash attention -> proj + residual
        h3 = self._scratch("h3", [n, ho, wo, cout])
        h3_lo = self._scratch("h3_lo", [n, ho, wo, cout]) if self.x3 else None
        self.gn(mid, b.norm2, pk[key + ".norm2"], h3, silu=False, y_lo=h3_lo, stats=st2["stats"] if st2 else None)
        f32_attn = self._f32_attention(b)
        qkv = self._scratch("qkv_f32" if f32_attn else "qkv", [n, ho, wo, 3 * cout])
        wq = pk[key + ".qkv"]
        self.conv(self.segs3(h3, h3_lo, cout, wq["hi"], wq.get("lo"), cout, 1), n, ho, wo, 3 * cout, qkv, 3 * cout,
                  bias=wq["bias"])
        att = self._scratch("att", [n, ho, wo, cout])
        wp = pk[key + ".proj"]
        if f32_attn:
            # fp32 qkv (the three-segment GEMM above wrote it unrounded) -> fp32 softmax(q.k) v -> bf16 hi + lo
            att_lo = self._scratch("att_lo", [n, ho, wo, cout])
            self._launch(L.lib().fg_attention_f32, qkv.data_ptr(), 3 * cout, n, ho * wo, b.num_heads,
                         cout // b.num_heads, att.data_ptr(), att_lo.data_ptr(), cout,
                         flops=4.0 * n * (ho * wo) ** 2 * cout)
            segs = self.segs3(att, att_lo, cout, wp["hi"], wp["lo"], cout, 1)
        else:
            self._launch(L.lib().fg_attention, qkv.data_ptr(), 3 * cout, n, ho * wo, b.num_heads,
                         cout // b.num_heads, att.data_ptr(), cout, flops=4.0 * n * (ho * wo) ** 2 * cout)
            segs = [ops.Seg(att, cout, wp["hi"], cout, 1)]
            if self.x3:
                segs.append(ops.Seg(att, cout, wp["lo"], cout, 1, alg=False))
        self.conv(segs, n, ho, wo, cout, dst.t, dst.ld, bias=wp["bias"], res=mid.t, res_ld=mid.ld,
                  res_mode=L.FG_RES_SAME, out_scale=float(b.skip_scale), **(out_stats or {}))

    def run(self, stream: int) -> None:
        check = L.check
        for fn, args in self.steps:
            st = fn(*args, stream)
            if st:
                check(st, fn.__name__)

    def run_timed(self, stream: int) -> Dict[str, Dict[str, float]]:
        """Same launches with a CUDA event pair around each one (measurement pass of bench.py): returns, per entry
        point, the launch count, the summed device time in ms and the summed algorithmic FLOPs."""
        evs = []
        for (fn, args), fl in zip(self.steps, self.step_flops):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            st = fn(*args, stream)
            b.record()
            if st:
                L.check(st, fn.__name__)
            evs.append((fn.__name__, a, b, fl))
        torch.cuda.synchronize()
        out: Dict[str, Dict[str, float]] = {}
        detail = getattr(self.net, "profile_detail", None)
        for (name, a, b, fl), (desc, nbytes) in zip(evs, self.step_desc):
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0})
            ms = a.elapsed_time(b)
            d["launches"] += 1
            d["ms"] += ms
            d["flops"] += fl
            if detail is not None:
                detail.append((name, desc, ms, fl, nbytes))
        return out


# ---------------------------------------------------------------------------------------------------------
# the network
# ---------------------------------------------------------------------------------------------------------
class EDMPrecond(FastGenNetwork):
    """B200 drop-in for ``fastgen.networks.EDM.network.EDMPrecond``."""

    def __init__(self, img_resolution, img_channels, label_dim=0, sigma_data=0.5, sigma_shift=0.0,
                 model_type="DhariwalUNet", drop_precond=None, net_pred_type="x0", schedule_type="edm",
                 compute: str = "bf16", **model_kwargs):
        super().__init__(net_pred_type=net_pred_type, schedule_type=schedule_type, **model_kwargs)
        if drop_precond is not None and drop_precond not in ["input", "output", "both"]:
            raise ValueError(f"drop_precond must be one of 'input', 'output', 'both', or None, got {drop_precond}")
        if compute not in ("bf16", "bf16x3"):
            raise ValueError(f"compute must be 'bf16' or 'bf16x3', got {compute}")
        self.img_resolution, self.img_channels, self.label_dim = img_resolution, img_channels, label_dim
        self.sigma_data, self.sigma_shift, self.drop_precond = sigma_data, sigma_shift, drop_precond
        self.compute = compute
        self.fuse_gn_stats = True   # GroupNorm statistics accumulated by the conv that writes the tensor
        self.bf16_h1 = os.environ.get("FASTGEN_B200_BF16_H1", "1") != "0"   # conv0 -> norm1 tensor in bf16 (compute="bf16")
        # attention of the "bf16x3" path: "bf16" = the tensor-core kernel (q, k, v, softmax weights rounded to bf16:
        # ~1e-3 of the fp32 reference), "f32" = fg_attention_f32 (CUDA cores; profiles/x3_emulation.py predicts ~1e-5)
        if _X3_ATTENTION_DEFAULT not in ("f32", "bf16"):
            raise ValueError(f"FASTGEN_B200_X3_ATTENTION must be 'f32' or 'bf16', got {_X3_ATTENTION_DEFAULT}")
        self.x3_attention = _X3_ATTENTION_DEFAULT
        self.model = _UNet(model_type, img_resolution=img_resolution, in_channels=img_channels,
                           out_channels=img_channels, label_dim=label_dim, **model_kwargs)
        self._packed: Optional[Dict[str, Any]] = None
        self._plans: Dict[int, _Plan] = {}
        self.profile: Optional[Dict[str, Dict[str, float]]] = None  # set to {} to time every launch (bench.py)
        self.use_cuda_graphs = False   # opt-in: replay a batch size's launch list as one CUDA graph (see raw_forward)
        self.graph_launches = 0        # kernel launches replayed through graphs (bench bookkeeping)

    # ---- weight packing ---------------------------------------------------------------------------------
    def _invalidate(self):
        self._packed = None
        self._plans = {}

    def _apply(self, fn, *a, **kw):
        self._invalidate()
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._invalidate()
        return super().load_state_dict(*a, **kw)

    def set_compute(self, compute: str) -> None:
        if compute not in ("bf16", "bf16x3"):
            raise ValueError(compute)
        if compute != self.compute:
            self.compute = compute
            self._invalidate()

    def set_x3_attention(self, mode: str) -> None:
        """attention of the "bf16x3" path: "f32" (fg_attention_f32, default) or "bf16" (the tensor-core kernel)"""
        if mode not in ("f32", "bf16"):
            raise ValueError(f"x3_attention must be 'f32' or 'bf16', got {mode}")
        if mode != self.x3_attention:
            self.x3_attention = mode
            self._plans = {}

    @torch.no_grad()
    def repack(self) -> None:
        """Convert the fp32 master parameters into kernel layouts (bf16 [cout][taps][cin], hi/lo splits, ...)."""
        m = self.model
        dev = next(self.parameters()).device
        if dev.type != "cuda" and not _ALLOW_HOST_PLANS:
            raise L.FastGenB200Error("EDMPrecond (B200) must live on a CUDA device: there is no CPU path")
        L.lib()  # fail loudly if the library is missing
        pk: Dict[str, Any] = {}

        def conv_w(c: _Conv, cin_pad: Optional[int] = None, perm: Optional[torch.Tensor] = None, scale: float = 1.0):
            w = c.weight.detach().float() * scale                      # [cout, cin, k, k]
            b = c.bias.detach().float() if c.bias is not None else torch.zeros(w.shape[0], device=dev)
            if perm is not None:
                w, b = w[perm], b[perm]
            cout, cin, k, _ = w.shape
            w = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)      # [cout][tap][cin], tap = kh*3+kw
            if cin_pad is not None and cin_pad != cin:
                w = torch.nn.functional.pad(w, (0, cin_pad - cin))
            hi, lo = _split(w.contiguous())
            return {"hi": hi, "lo": lo, "bias": b.contiguous()}

        def lin_w(l: _Linear, k_pad: Optional[int] = None, scale: float = 1.0):
            w = l.weight.detach().float() * scale
            if k_pad is not None and k_pad != w.shape[1]:
                w = torch.nn.functional.pad(w, (0, k_pad - w.shape[1]))
            hi, lo = _split(w.contiguous())
            d = {"hi": hi, "lo": lo}
            if l.bias is not None:
                d["bias"] = l.bias.detach().float().contiguous()
            return d

        def gn_w(g: _GroupNorm):
            return {"gamma": g.weight.detach().float().contiguous(), "beta": g.bias.detach().float().contiguous()}

        song = m.model_type == "SongUNet"
        if m.map_label is not None:
            pk["map_label"] = lin_w(m.map_label, _pad8(m.label_dim), math.sqrt(m.label_dim) if song else 1.0)
        if m.map_augment is not None:
            pk["map_augment"] = lin_w(m.map_augment, _pad8(m.augment_dim))
        pk["map_layer0"] = lin_w(m.map_layer0)
        pk["map_layer1"] = lin_w(m.map_layer1)

        aff_w, aff_b, off = [], [], 0
        for side, md in (("enc", m.enc), ("dec", m.dec)):
            for name, mod in md.items():
                key = f"{side}.{name}"
                if isinstance(mod, _Conv):
                    if "aux" in name:
                        continue
                    pk[key] = conv_w(mod, _pad8(mod.in_channels))
                elif isinstance(mod, _Block):
                    pk[key + ".norm0"] = gn_w(mod.norm0)
                    pk[key + ".conv0"] = conv_w(mod.conv0)
                    pk[key + ".norm1"] = gn_w(mod.norm1)
                    pk[key + ".conv1"] = conv_w(mod.conv1)
                    pk[key + ".affine_off"] = off
                    aff_w.append(mod.affine.weight.detach().float())
                    aff_b.append(mod.affine.bias.detach().float())
                    off += mod.affine.out_features
                    # keep every block's slice 4-aligned for the vectorised epilogue
                    pad = (-off) % 4
                    if pad:
                        aff_w.append(torch.zeros(pad, mod.emb_channels, device=dev))
                        aff_b.append(torch.zeros(pad, device=dev))
                        off += pad
                    if mod.skip is not None and mod.skip.kernel == 1:
                        pk[key + ".skip"] = conv_w(mod.skip)
                        pk[key + ".bias1s"] = (pk[key + ".conv1"]["bias"] + pk[key + ".skip"]["bias"]).contiguous()
                    if mod.num_heads:
                        c, hds = mod.out_channels, mod.num_heads
                        # reference channel o = (h*dh + c)*3 + j  ->  ours j*C + h*dh + c
                        perm = torch.arange(3 * c, device=dev).reshape(c, 3).t().reshape(-1)
                        pk[key + ".norm2"] = gn_w(mod.norm2)
                        pk[key + ".qkv"] = conv_w(mod.qkv, perm=perm)
                        pk[key + ".proj"] = conv_w(mod.proj)
        w_all = torch.cat(aff_w, 0).contiguous()
        hi, lo = _split(w_all)
        pk["affine_all"] = {"hi": hi, "lo": lo, "bias": torch.cat(aff_b, 0).contiguous()}
        norm, conv = m.final_layers()
        pk["final.norm"] = gn_w(norm)
        pk["final.conv"] = conv_w(conv)
        self._packed = pk
        self._plans = {}

    def _plan(self, n: int) -> _Plan:
        if self._packed is None:
            self.repack()
        p = self._plans.get(n)
        if p is None:
            p = _Plan(self, n)
            self._plans[n] = p
        return p

    # ---- forward ------------------------------------------------------------------------------------------
    def _coefficients(self, t: torch.Tensor):
        """fp64 preconditioning coefficients, same formulas as the reference (EDM/network.py:771-805)."""
        t = t.to(torch.float64).reshape(-1)
        sd = self.sigma_data
        c_in = 1 / (sd ** 2 + t ** 2).sqrt()
        ts = t - self.sigma_shift if (not self.training and self.sigma_shift is not None) else t
        c_skip = sd ** 2 / (ts ** 2 + sd ** 2)
        c_out = ts * sd / (ts ** 2 + sd ** 2).sqrt()
        return c_in.contiguous(), c_skip.contiguous(), c_out.contiguous()

    def raw_forward(self, x_t: torch.Tensor, t: torch.Tensor, condition: Any = None,
                    r: Optional[torch.Tensor] = None) -> Tuple[_Plan, torch.Tensor, torch.Tensor]:
        """Run the UNet; the raw output F stays channels-last in ``plan.f_out``.  Returns (plan, c_skip, c_out)."""
        if not x_t.is_cuda:
            raise L.FastGenB200Error("EDMPrecond (B200) needs CUDA tensors: there is no CPU path")
        if x_t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"x_t must be float32 or bfloat16, got {x_t.dtype}")
        m = self.model
        n = x_t.shape[0]
        if tuple(x_t.shape[1:]) != (m.in_channels, m.img_resolution, m.img_resolution):
            raise ValueError(f"x_t shape {tuple(x_t.shape)} does not match the network")
        plan = self._plan(n)
        x_t = x_t.contiguous()
        t = torch.atleast_1d(t).to(x_t.device)
        if t.numel() == 1 and n > 1:
            t = t.expand(n)
        c_in, c_skip, c_out = self._coefficients(t)
        eps = self.noise_scheduler.clamp_min
        if self.drop_precond in ("input", "both"):
            plan.t_noise.copy_(t.to(x_t.dtype))
            ops.edm_input(x_t, None, plan.xin, plan.xin.shape[-1], xin_lo=plan.xin_lo)
            if r is not None:
                r_noise = r
        else:
            plan.t_noise.copy_((t.clamp(min=eps).log() / 4).to(x_t.dtype))
            ops.edm_input(x_t, c_in, plan.xin, plan.xin.shape[-1], xin_lo=plan.xin_lo)
            if r is not None:
                r_noise = r.clamp(min=eps).log() / 4
        if m.has_r:
            if r is None:
                raise ValueError("r_noise_labels provided, but r_timestep is not set")
            plan.r_noise.copy_(torch.atleast_1d(r_noise).expand(n).to(x_t.dtype))
        elif r is not None and m.model_type == "SongUNet":
            raise ValueError("r_noise_labels provided, but r_timestep is not set")
        # conditioning (reference :903-925)
        augment = None
        if isinstance(condition, dict) and "aug_condition" in condition:
            augment = condition["aug_condition"]
            condition = condition["orig_condition"]
            if m.map_augment is None or augment.shape[-1] != m.augment_dim:
                augment = None
        if m.label_dim:
            if condition is None:
                plan.labels.zero_()
            else:
                lab = condition.reshape(-1, m.label_dim).to(device=x_t.device, dtype=torch.float32)
                plan.labels[:, :m.label_dim].copy_(lab.expand(n, m.label_dim))
        if plan.augment is not None:
            if augment is None:
                plan.augment.zero_()
            else:
                plan.augment[:, :m.augment_dim].copy_(augment.reshape(-1, m.augment_dim).to(torch.float32).expand(n, -1))
        stream = torch.cuda.current_stream().cuda_stream
        if self.profile is None:
            # opt-in (use_cuda_graphs): replay the recorded launch list as ONE CUDA graph from its second use on.  The plan
            # reads only its own static buffers (the inputs were staged into them above), so the captured graph is valid
            # for every later call at this batch size.  Pays off when the step is host-launch bound (CIFAR-10 batch 64:
            # 174 launches of ~10-30 us); the ImageNet-64 batch-256 step is GPU bound and gains nothing.
            if self.use_cuda_graphs and plan.graph is None and plan.runs >= 1:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    plan.run(torch.cuda.current_stream().cuda_stream)
                plan.graph = g
            if self.use_cuda_graphs and plan.graph is not None:
                plan.graph.replay()
                self.graph_launches += len(plan.steps)
            else:
                plan.run(stream)
            plan.runs += 1
        else:  # measurement pass: per entry point launch count / device ms / algorithmic FLOPs, accumulated
            for name, d in plan.run_timed(stream).items():
                acc = self.profile.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0})
                for k in acc:
                    acc[k] += d[k]
        return plan, c_skip, c_out

    def forward(self, x_t: torch.Tensor, t: torch.Tensor, condition: Any = None, r: Optional[torch.Tensor] = None,
                return_features_early: bool = False, feature_indices: Optional[Set[int]] = None,
                return_logvar: bool = False, fwd_pred_type: Optional[str] = None, **fwd_kwargs):
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
        with torch.no_grad():
            plan, c_skip, c_out = self.raw_forward(x_t, t, condition, r)
            x_t = x_t.contiguous()
            out = torch.empty_like(x_t)
            if self.drop_precond in ("output", "both"):
                ops.edm_step(plan.f_out, plan.f_ld, x_t, out, None, None)
            else:
                ops.edm_step(plan.f_out, plan.f_ld, x_t, out, c_skip, c_out)
            tt = torch.atleast_1d(t).to(x_t.device)
            if tt.numel() == 1 and x_t.shape[0] > 1:
                tt = tt.expand(x_t.shape[0])
            return self.noise_scheduler.convert_model_output(x_t, out, tt, src_pred_type=self.net_pred_type,
                                                             target_pred_type=fwd_pred_type)

    def sample(self, noise: torch.Tensor, condition: Optional[torch.Tensor] = None,
               neg_condition: Optional[torch.Tensor] = None, guidance_scale: Optional[float] = 5.0, num_steps: int = 50,
               **kwargs) -> torch.Tensor:
        """Teacher sampling: EDM's deterministic Euler sampler with optional classifier-free guidance -- reference
        ``EDMPrecond.sample`` (EDM/network.py:976-1026): same sigma list, same batched [negative | positive] network call,
        same update ``x <- x + (sigma_next - sigma) (x - x0) / sigma``.  The two elementwise passes of a step are
        ``fg_axpby`` launches with fp64 coefficients and one rounding: ``x0 = (1 - g) x0_uncond + g x0_cond`` and
        ``x <- (sigma_next / sigma) x + (1 - sigma_next / sigma) x0`` (the reference rounds after each fp32 operation)."""
        assert self.schedule_type == "edm", f"{self.schedule_type} is not supported"
        with torch.no_grad():
            sched = self.noise_scheduler
            sigmas = sched.get_t_list(num_steps, device=noise.device)
            x = sched.latents(noise=noise, t_init=sigmas[0])
            B = x.shape[0]
            cfg = guidance_scale is not None and guidance_scale > 1.0 and neg_condition is not None
            if cfg:
                cond_input = torch.cat([neg_condition, condition], dim=0)
                g = torch.full((B,), float(guidance_scale), dtype=torch.float64, device=x.device)
            for sigma, sigma_next in zip(sigmas[:-1], sigmas[1:]):
                t = sigma.expand(B)
                if cfg:
                    x0_pred = self(torch.cat([x, x], dim=0), torch.cat([t, t], dim=0), condition=cond_input,
                                   fwd_pred_type="x0")
                    x0_uncond, x0_cond = x0_pred.chunk(2)
                    x0_pred = ops.axpby(x0_uncond.contiguous(), 1.0 - g, x0_cond.contiguous(), g)
                else:
                    x0_pred = self(x, t, condition=condition, fwd_pred_type="x0")
                ratio = (sigma_next.to(torch.float64) / sigma.to(torch.float64)).expand(B).contiguous()
                x = ops.axpby(x.contiguous(), ratio, x0_pred.contiguous(), 1.0 - ratio)
            return x


EDMPrecondB200 = EDMPrecond
