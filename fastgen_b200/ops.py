"""Thin torch-tensor wrappers over the C ABI (``include/fastgen_b200.h``).

PyTorch is used for device memory and streams only: every function here passes raw device
pointers and sizes to ``libfastgen_b200.so`` on the current CUDA stream (so the calls are
CUDA-graph capturable) and raises if the library reports an error.  No function has a
PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

_DT = {torch.float32: L.FG_F32, torch.bfloat16: L.FG_BF16, torch.float64: L.FG_F64}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.FastGenB200Error("fastgen_b200 kernels need CUDA tensors (there is no CPU path)")


class Seg:
    """One K segment of :func:`conv_gemm`: bf16 activation view [N,H,W,cin] (pixel stride ``ld``) and its weight."""

    __slots__ = ("a", "ld", "w", "cin", "taps", "alg")

    def __init__(self, a: torch.Tensor, ld: int, w: torch.Tensor, cin: int, taps: int, alg: bool = True):
        self.a, self.ld, self.w, self.cin, self.taps = a, ld, w, cin, taps
        self.alg = alg  # False for the extra hi*lo / lo*hi products of split precision (not algorithmic FLOPs)


def conv_gemm_params(
    segs: Sequence[Seg],
    n: int, h: int, w: int, cout: int,
    out: torch.Tensor, out_ld: int,
    bias: Optional[torch.Tensor] = None,
    rowvec: Optional[torch.Tensor] = None, rowvec_ld: int = 0,
    res: Optional[torch.Tensor] = None, res_ld: int = 0, res_mode: int = L.FG_RES_NONE,
    out_scale: float = 1.0,
    act: int = L.FG_ACT_NONE,
    gate: Optional[torch.Tensor] = None, gate_ld: int = 0, gate_rows: int = 0,
    stats: Optional[torch.Tensor] = None, stats_groups: int = 0, stats_cg: int = 0, stats_ch0: int = 0,
    stats2: Optional[torch.Tensor] = None, stats2_groups: int = 0, stats2_cg: int = 0, stats2_ch0: int = 0,
) -> "L.ConvParams":
    """Fill an ``fg_conv_params`` (the tensors must stay alive and in place while the struct is used)."""
    _need_cuda(out, bias, rowvec, res, gate, *[s.a for s in segs], *[s.w for s in segs])
    p = L.ConvParams()
    p.n, p.h, p.w, p.cout, p.nseg = n, h, w, cout, len(segs)
    if len(segs) > L.FG_MAX_SEG:
        raise ValueError(f"conv_gemm: {len(segs)} segments > FG_MAX_SEG={L.FG_MAX_SEG}")
    for i, s in enumerate(segs):
        if s.a.dtype != torch.bfloat16 or s.w.dtype != torch.bfloat16:
            raise ValueError("conv_gemm: operands must be bfloat16")
        p.seg[i].a, p.seg[i].a_ld, p.seg[i].w, p.seg[i].cin, p.seg[i].taps = s.a.data_ptr(), s.ld, s.w.data_ptr(), s.cin, s.taps
    p.bias = _ptr(bias)
    p.rowvec, p.rowvec_ld = _ptr(rowvec), rowvec_ld
    p.act = act
    p.gate, p.gate_ld, p.gate_rows = _ptr(gate), gate_ld, gate_rows
    p.res, p.res_ld, p.res_mode = _ptr(res), res_ld, (res_mode if res is not None else L.FG_RES_NONE)
    p.res_dtype = _DT[res.dtype] if res is not None else 0
    p.out_scale = out_scale
    p.out, p.out_dtype, p.out_ld = out.data_ptr(), _DT[out.dtype], out_ld
    if stats is not None:
        _need_cuda(stats)
        assert stats.dtype == torch.float64
        p.stats, p.stats_groups, p.stats_cg, p.stats_ch0 = stats.data_ptr(), stats_groups, stats_cg, stats_ch0
    if stats2 is not None:
        _need_cuda(stats2)
        assert stats is not None and stats2.dtype == torch.float64
        p.stats2, p.stats2_groups, p.stats2_cg, p.stats2_ch0 = stats2.data_ptr(), stats2_groups, stats2_cg, stats2_ch0
    return p


def conv_gemm(*args, **kwargs) -> None:
    """out = out_scale * (sum_s conv(A_s, W_s) + bias + rowvec[n] + resample(res)) -- see ``fg_conv_gemm``."""
    p = conv_gemm_params(*args, **kwargs)
    L.check(L.lib().fg_conv_gemm(C.byref(p), _stream()), "fg_conv_gemm")


def gn_stats(x: torch.Tensor, x_ld: int, n: int, hw: int, c: int, groups: int, stats: torch.Tensor) -> None:
    _need_cuda(x, stats)
    assert stats.dtype == torch.float64 and stats.numel() >= 2 * n * groups
    L.check(L.lib().fg_gn_stats(x.data_ptr(), _DT[x.dtype], x_ld, n, hw, c, groups, stats.data_ptr(), _stream()),
            "fg_gn_stats")


def gn_apply_params(
    x: torch.Tensor, x_ld: int, n: int, h: int, w: int, c: int, groups: int,
    stats: torch.Tensor, eps: float, gamma: torch.Tensor, beta: torch.Tensor,
    y: torch.Tensor, y_ld: int,
    scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None, mod_ld: int = 0,
    silu: bool = True, resample: int = L.FG_RES_NONE,
    y_lo: Optional[torch.Tensor] = None,
    raw: Optional[torch.Tensor] = None, raw_ld: int = 0, raw_lo: Optional[torch.Tensor] = None,
    raw_f32: Optional[torch.Tensor] = None, raw_f32_ld: int = 0,
) -> "L.GnApplyParams":
    _need_cuda(x, stats, gamma, beta, y, scale, shift, y_lo, raw, raw_lo, raw_f32)
    p = L.GnApplyParams()
    p.x, p.x_dtype, p.x_ld = x.data_ptr(), _DT[x.dtype], x_ld
    p.n, p.h, p.w, p.c, p.groups = n, h, w, c, groups
    p.stats, p.eps = stats.data_ptr(), eps
    p.gamma, p.beta = gamma.data_ptr(), beta.data_ptr()
    p.scale, p.shift, p.mod_ld = _ptr(scale), _ptr(shift), mod_ld
    p.silu, p.resample = int(silu), resample
    p.y, p.y_ld, p.y_lo = y.data_ptr(), y_ld, _ptr(y_lo)
    p.raw, p.raw_ld, p.raw_lo = _ptr(raw), raw_ld, _ptr(raw_lo)
    p.raw_f32, p.raw_f32_ld = _ptr(raw_f32), raw_f32_ld
    return p


def gn_apply(*args, **kwargs) -> None:
    p = gn_apply_params(*args, **kwargs)
    L.check(L.lib().fg_gn_apply(C.byref(p), _stream()), "fg_gn_apply")


def attention(qkv: torch.Tensor, qkv_ld: int, n: int, tokens: int, heads: int, d: int,
              out: torch.Tensor, out_ld: int) -> None:
    _need_cuda(qkv, out)
    L.check(L.lib().fg_attention(qkv.data_ptr(), qkv_ld, n, tokens, heads, d, out.data_ptr(), out_ld, _stream()),
            "fg_attention")


class PreparedConv:
    """A validated, tile-selected, descriptor-encoded fg_conv_gemm launch (``fg_conv_gemm_prepare``): ``fn(*args, stream)``
    replays it.  Holds the 64-byte aligned host blob the library filled; the caller keeps the tensors alive."""

    __slots__ = ("buf", "ptr", "fn", "args")

    def __init__(self, p: "L.ConvParams"):
        lib = L.lib()
        size = int(lib.fg_conv_gemm_prepared_size())
        self.buf = C.create_string_buffer(size + 64)
        self.ptr = (C.addressof(self.buf) + 63) & ~63
        L.check(lib.fg_conv_gemm_prepare(C.byref(p), self.ptr, size), "fg_conv_gemm")
        self.fn, self.args = lib.fg_conv_gemm_run, (self.ptr,)


def conv_gemm_launch(p: "L.ConvParams"):
    """(fn, args, keep) of a launch-list entry for the convolution ``p``: the prepared form on a CUDA machine; the plain
    ``fg_conv_gemm(params)`` call where there is no device to encode TMA descriptors for (host-side plan validation in the
    CPU tests, which expects the call itself to fail with FG_ERR_CUDA)."""
    if torch.cuda.is_available():
        pc = PreparedConv(p)
        return pc.fn, pc.args, pc
    return L.lib().fg_conv_gemm, (C.byref(p),), p


def sampler_vec(enable: Optional[bool] = None) -> bool:
    """Select (True / False) or query (None) the vectorised forms of the sampler-step kernels; returns the previous
    setting.  Default on (FASTGEN_B200_SAMPLER_VEC=0 turns them off): see ``fg_sampler_vec``."""
    return bool(L.lib().fg_sampler_vec(-1 if enable is None else int(bool(enable))))


def attention_f32(qkv: torch.Tensor, qkv_ld: int, n: int, tokens: int, heads: int, d: int,
                  out_hi: torch.Tensor, out_lo: Optional[torch.Tensor], out_ld: int) -> None:
    """fp32 softmax(q.k / sqrt(d)) v of a packed fp32 qkv tensor, written as a bf16 hi (+ lo) pair."""
    _need_cuda(qkv, out_hi, out_lo)
    if qkv.dtype != torch.float32 or out_hi.dtype != torch.bfloat16 or (out_lo is not None and out_lo.dtype != torch.bfloat16):
        raise ValueError("attention_f32: qkv must be float32, the outputs bfloat16")
    L.check(L.lib().fg_attention_f32(qkv.data_ptr(), qkv_ld, n, tokens, heads, d, out_hi.data_ptr(), _ptr(out_lo), out_ld,
                                     _stream()), "fg_attention_f32")


def edm_input(x: torch.Tensor, c_in: Optional[torch.Tensor], xin: torch.Tensor, c_pad: int,
              xin_lo: Optional[torch.Tensor] = None) -> None:
    _need_cuda(x, c_in, xin, xin_lo)
    n, c = x.shape[0], x.shape[1]
    hw = x[0, 0].numel()
    assert x.is_contiguous()
    L.check(L.lib().fg_edm_input(x.data_ptr(), _DT[x.dtype], n, c, hw, _ptr(c_in), xin.data_ptr(), _ptr(xin_lo), c_pad,
                                 _stream()), "fg_edm_input")


def edm_step(
    f: torch.Tensor, f_ld: int, x_t: torch.Tensor, x0: torch.Tensor,
    c_skip: Optional[torch.Tensor], c_out: Optional[torch.Tensor],
    x_next: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None,
    alpha_next: Optional[torch.Tensor] = None, sigma_next: Optional[torch.Tensor] = None,
    ode: bool = False, alpha_cur: Optional[torch.Tensor] = None, sigma_cur: Optional[torch.Tensor] = None,
) -> None:
    _need_cuda(f, x_t, x0, c_skip, c_out, x_next, eps, alpha_next, sigma_next, alpha_cur, sigma_cur)
    assert x_t.is_contiguous() and x0.is_contiguous() and f.dtype == torch.float32
    p = L.EdmStepParams()
    p.f, p.f_ld = f.data_ptr(), f_ld
    p.x_t, p.eps, p.x_dtype = x_t.data_ptr(), _ptr(eps), _DT[x_t.dtype]
    p.n, p.c, p.hw = x_t.shape[0], x_t.shape[1], x_t[0, 0].numel()
    p.c_skip, p.c_out = _ptr(c_skip), _ptr(c_out)
    p.ode = int(ode)
    p.alpha_cur, p.sigma_cur = _ptr(alpha_cur), _ptr(sigma_cur)
    p.alpha_next, p.sigma_next = _ptr(alpha_next), _ptr(sigma_next)
    p.x0, p.x_next = x0.data_ptr(), _ptr(x_next)
    L.check(L.lib().fg_edm_step(C.byref(p), _stream()), "fg_edm_step")


def _coef_extents(a: torch.Tensor, coef: torch.Tensor) -> Tuple[int, int, int, int]:
    """(total, outer, mid, inner) for a coefficient tensor that broadcasts against ``a`` like the reference's
    ``expand_like``: a scalar, one value per sample ([B]), or one per (sample, frame) ([B,1,T,1,1] for 5-D ``a``)."""
    total = a.numel()
    k = coef.numel()
    b = a.shape[0]
    if k == 1:
        return total, total, 1, total
    if k == b:
        return total, total // b, 1, total // b
    if a.dim() == 5 and k == b * a.shape[2]:
        hw = a.shape[3] * a.shape[4]
        return total, total // b, a.shape[2], hw
    raise ValueError(f"coefficient of {k} elements does not broadcast against tensor {tuple(a.shape)}")


def axpby(a: torch.Tensor, ca: torch.Tensor, b: Optional[torch.Tensor] = None, cb: Optional[torch.Tensor] = None,
          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = cast(double(a)*ca + double(b)*cb) with fp64 coefficients broadcast per sample / per frame (one rounding)."""
    _need_cuda(a, b, ca, cb)
    assert a.is_contiguous() and (b is None or b.is_contiguous())
    assert ca.dtype == torch.float64 and (cb is None or cb.dtype == torch.float64)
    out = torch.empty_like(a) if out is None else out
    total, outer, mid, inner = _coef_extents(a, ca)
    L.check(L.lib().fg_axpby(a.data_ptr(), _ptr(b), _DT[a.dtype], total, outer, mid, inner, ca.data_ptr(), _ptr(cb),
                             out.data_ptr(), _stream()), "fg_axpby")
    return out


def sub_div(a: torch.Tensor, b: torch.Tensor, cb: torch.Tensor, cd: torch.Tensor,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = cast((double(a) - double(b)*cb) / cd)."""
    _need_cuda(a, b, cb, cd)
    assert a.is_contiguous() and b.is_contiguous()
    out = torch.empty_like(a) if out is None else out
    total, outer, mid, inner = _coef_extents(a, cb)
    L.check(L.lib().fg_sub_div(a.data_ptr(), b.data_ptr(), _DT[a.dtype], total, outer, mid, inner, cb.data_ptr(),
                               cd.data_ptr(), out.data_ptr(), _stream()), "fg_sub_div")
    return out


def pos_embedding(t: torch.Tensor, channels: int, max_positions: int, endpoint: bool, swap: bool,
                  out: torch.Tensor, out_ld: int) -> None:
    _need_cuda(t, out)
    assert t.dtype == torch.float32 and out.dtype == torch.float32
    L.check(L.lib().fg_pos_embedding(t.data_ptr(), t.numel(), channels, max_positions, int(endpoint), int(swap),
                                     out.data_ptr(), out_ld, _stream()), "fg_pos_embedding")


def split_bf16(x: torch.Tensor, x_ld: int, n: int, c: int, hi: torch.Tensor, out_ld: int,
               lo: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None, add_ld: int = 0,
               act: int = 0, y: Optional[torch.Tensor] = None, y_ld: int = 0) -> None:
    _need_cuda(x, hi, lo, add, y)
    L.check(L.lib().fg_split_bf16(x.data_ptr(), x_ld, _ptr(add), add_ld, n, c, act, hi.data_ptr(), _ptr(lo), out_ld,
                                  _ptr(y), y_ld, _stream()), "fg_split_bf16")


def launch_count() -> int:
    return int(L.lib().fg_launch_count())


def launch_count_reset() -> None:
    L.lib().fg_launch_count_reset()


# ---------------------------------------------------------------------------------------------------------
# transformer (Wan / DiT) kernels
# ---------------------------------------------------------------------------------------------------------
def flash_attn_params(q: torch.Tensor, q_ld: int, q_bs: int, segs: Sequence[Tuple[torch.Tensor, torch.Tensor, int, int, int]],
                      batch: int, heads: int, d: int, tq: int, out: torch.Tensor, out_ld: int, out_bs: int,
                      scale: float = 0.0, mask_first: int = 0, mask_chunk: int = 0,
                      workspace: Optional[torch.Tensor] = None,
                      kv_len0_per_sample: Optional[torch.Tensor] = None) -> "L.FlashAttnParams":
    """``segs``: up to two key/value segments ``(k, v, ld, batch_stride, length)`` (cache prefix, current chunk)."""
    if not 1 <= len(segs) <= 2:
        raise ValueError("flash_attn: one or two key/value segments")
    _need_cuda(q, out, *[s[0] for s in segs], *[s[1] for s in segs])
    p = L.FlashAttnParams()
    p.q, p.q_ld, p.q_bs = q.data_ptr(), q_ld, q_bs
    for i, (k, v, ld, bs, ln) in enumerate(segs):
        if k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
            raise ValueError("flash_attn: operands must be bfloat16")
        p.k[i], p.v[i], p.kv_ld[i], p.kv_bs[i], p.kv_len[i] = k.data_ptr(), v.data_ptr(), ld, bs, ln
    p.batch, p.heads, p.d, p.tq = batch, heads, d, tq
    p.scale, p.mask_first, p.mask_chunk = scale, mask_first, mask_chunk
    p.out, p.out_ld, p.out_bs = out.data_ptr(), out_ld, out_bs
    if workspace is not None:
        _need_cuda(workspace)
        p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    if kv_len0_per_sample is not None:   # int32 [batch] on the device: a prefix attention mask over segment 0
        _need_cuda(kv_len0_per_sample)
        if kv_len0_per_sample.dtype != torch.int32 or kv_len0_per_sample.numel() != batch or not kv_len0_per_sample.is_contiguous():
            raise ValueError("flash_attn: kv_len0_per_sample must be a contiguous int32 tensor of `batch` elements")
        p.kv_len0_per_sample = kv_len0_per_sample.data_ptr()
    return p


def flash_attn_workspace_size(p: "L.FlashAttnParams") -> int:
    return int(L.lib().fg_flash_attn_workspace_size(C.byref(p)))


def flash_attn(*args, **kwargs) -> None:
    p = flash_attn_params(*args, **kwargs)
    L.check(L.lib().fg_flash_attn(C.byref(p), _stream()), "fg_flash_attn")


def ln_mod_params(x: torch.Tensor, x_ld: int, rows: int, d: int, eps: float, y: torch.Tensor, y_ld: int,
                  gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None,
                  scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None, mod_ld: int = 0,
                  rows_per_group: int = 0) -> "L.LnModParams":
    _need_cuda(x, y, gamma, beta, scale, shift)
    if x.dtype not in (torch.bfloat16, torch.float32) or y.dtype != torch.bfloat16:
        raise ValueError("ln_modulate: input rows must be bfloat16 or float32, output rows bfloat16")
    p = L.LnModParams()
    p.x, p.x_ld, p.rows, p.d, p.eps = x.data_ptr(), x_ld, rows, d, eps
    p.x_dtype = _DT[x.dtype]
    p.gamma, p.beta, p.scale, p.shift = _ptr(gamma), _ptr(beta), _ptr(scale), _ptr(shift)
    p.mod_ld, p.rows_per_group = mod_ld, rows_per_group
    p.y, p.y_ld = y.data_ptr(), y_ld
    return p


def ln_modulate(*args, **kwargs) -> None:
    p = ln_mod_params(*args, **kwargs)
    L.check(L.lib().fg_ln_modulate(C.byref(p), _stream()), "fg_ln_modulate")


def qk_norm_rope_params(q: Optional[torch.Tensor], k: Optional[torch.Tensor], q_ld: int, k_ld: int, rows: int,
                        heads: int, d: int, eps: float, wq: Optional[torch.Tensor], wk: Optional[torch.Tensor],
                        rows_per_batch: int,
                        rope: Optional[Tuple[torch.Tensor, torch.Tensor, int, int]] = None,
                        grid: Tuple[int, int, int] = (1, 1, 1), frame0: int = 0, frame_rows: int = 0, tok0: int = 0,
                        v: Optional[torch.Tensor] = None, v_ld: int = 0,
                        k_cache: Optional[torch.Tensor] = None, v_cache: Optional[torch.Tensor] = None,
                        cache_ld: int = 0, cache_bs: int = 0, cache_row0: int = 0) -> "L.QkNormRopeParams":
    """``rope`` = (cos [max_pos, d/2] fp32, sin, nt, nh); ``grid`` = (frames, height, width) of the token grid."""
    _need_cuda(q, k, v, wq, wk, k_cache, v_cache, *(rope[:2] if rope else ()))
    p = L.QkNormRopeParams()
    p.q, p.k, p.v = _ptr(q), _ptr(k), _ptr(v)
    p.q_ld, p.k_ld, p.v_ld = q_ld, k_ld, v_ld
    p.rows, p.heads, p.d, p.eps = rows, heads, d, eps
    p.wq, p.wk = _ptr(wq), _ptr(wk)
    if rope is not None:
        cos, sin, nt, nh = rope
        assert cos.dtype == torch.float32 and sin.dtype == torch.float32 and cos.shape == sin.shape
        assert cos.shape[1] == d // 2 and cos.is_contiguous() and sin.is_contiguous()
        p.rope_cos, p.rope_sin, p.rope_max_pos, p.nt, p.nh = cos.data_ptr(), sin.data_ptr(), cos.shape[0], nt, nh
    p.rows_per_batch = rows_per_batch
    p.ppf, p.pph, p.ppw, p.frame0 = grid[0], grid[1], grid[2], frame0
    p.frame_rows, p.tok0 = frame_rows, tok0
    p.k_cache, p.v_cache, p.cache_ld, p.cache_bs, p.cache_row0 = _ptr(k_cache), _ptr(v_cache), cache_ld, cache_bs, cache_row0
    return p


def qk_norm_rope(*args, **kwargs) -> None:
    p = qk_norm_rope_params(*args, **kwargs)
    L.check(L.lib().fg_qk_norm_rope(C.byref(p), _stream()), "fg_qk_norm_rope")


def patchify(x: torch.Tensor, ph: int, pw: int, a: torch.Tensor, a_ld: int) -> None:
    _need_cuda(x, a)
    assert x.is_contiguous() and x.dim() == 5 and a.dtype == torch.bfloat16
    b, c, f, h, w = x.shape
    L.check(L.lib().fg_patchify(x.data_ptr(), _DT[x.dtype], b, c, f, h, w, ph, pw, a.data_ptr(), a_ld, _stream()),
            "fg_patchify")


def unpatchify(y: torch.Tensor, y_ld: int, out: torch.Tensor, ph: int, pw: int) -> None:
    _need_cuda(y, out)
    assert out.is_contiguous() and out.dim() == 5 and y.dtype == torch.bfloat16
    b, c, f, h, w = out.shape
    L.check(L.lib().fg_unpatchify(y.data_ptr(), y_ld, b, c, f, h, w, ph, pw, out.data_ptr(), _DT[out.dtype], _stream()),
            "fg_unpatchify")


def bcast_add(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> None:
    """out[i, g, :] = a[i, :] + b[g, :] (fp32)."""
    _need_cuda(a, b, out)
    assert a.dtype == b.dtype == out.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
    na, d = a.shape
    nb = b.shape[0]
    assert b.shape[1] == d and out.numel() == na * nb * d and out.is_contiguous()
    L.check(L.lib().fg_bcast_add(a.data_ptr(), b.data_ptr(), na, nb, d, out.data_ptr(), _stream()), "fg_bcast_add")


def strided_copy_args(src: torch.Tensor, dst: torch.Tensor, n: Sequence[int], src_strides: Sequence[int],
                      dst_strides: Sequence[int], item_elems: int):
    """ctypes arguments of ``fg_strided_copy`` for a 4-D index space; strides / item size in ELEMENTS of the tensors'
    (common) dtype.  Returns (args tuple, keep-alive objects)."""
    _need_cuda(src, dst)
    assert src.dtype == dst.dtype and len(n) == 4
    es = src.element_size()
    na = (C.c_int32 * 4)(*n)
    sa = (C.c_int64 * 4)(*[s * es for s in src_strides])
    da = (C.c_int64 * 4)(*[s * es for s in dst_strides])
    return (src.data_ptr(), dst.data_ptr(), na, sa, da, item_elems * es), (na, sa, da, src, dst)


def strided_copy(src, dst, n, src_strides, dst_strides, item_elems) -> None:
    args, _keep = strided_copy_args(src, dst, n, src_strides, dst_strides, item_elems)
    L.check(L.lib().fg_strided_copy(*args, _stream()), "fg_strided_copy")


def peer_put_params(src: torch.Tensor, desc, dst_ptrs: Sequence[int], flag_ptrs: Sequence[int], rank: int,
                    state: torch.Tensor, wait: bool = False, timeout_ns: int = 0) -> "L.PeerPutParams":
    """``fg_peer_put`` parameters from a ``parallel.peer_put_*_desc`` descriptor (strides / offsets in elements of
    ``src``'s dtype).  ``dst_ptrs[q]`` / ``flag_ptrs[q]``: device-mapped addresses of rank q's destination buffer and of
    its flag array for this channel (symmetric memory); ``state``: this channel's 4 local int64 words."""
    _need_cuda(src, state)
    P = len(dst_ptrs)
    if not 1 <= P <= L.FG_MAX_PEERS or len(flag_ptrs) != P or not 0 <= rank < P:
        raise ValueError(f"peer_put: {P} peers (max {L.FG_MAX_PEERS}), rank {rank}")
    if state.dtype != torch.int64 or state.numel() < 4 or not state.is_contiguous():
        raise ValueError("peer_put: state must be 4 contiguous int64 words")
    es = src.element_size()
    p = L.PeerPutParams()
    p.src = src.data_ptr()
    p.peers, p.rank = P, rank
    for q in range(P):
        p.dst[q], p.flags[q] = dst_ptrs[q], flag_ptrs[q]
    for i in range(4):
        p.n[i], p.src_strides[i], p.dst_strides[i] = desc["n"][i], desc["ss"][i] * es, desc["ds"][i] * es
    p.src_peer_stride, p.dst_offset, p.item_bytes = desc["src_peer_stride"] * es, desc["dst_off"] * es, desc["item"] * es
    p.state = state.data_ptr()
    if wait:    # fused wait on this rank's own flags of the channel: put + wait are one launch
        p.wait_flags, p.wait_timeout_ns = flag_ptrs[rank], timeout_ns
    return p


def geglu(h: torch.Tensor, rows: int, inner: int, y: torch.Tensor) -> None:
    """y[r, j] = h[r, j] * gelu(h[r, inner + j]) (erf GELU, rounded to bf16 before the product like diffusers' GEGLU)."""
    _need_cuda(h, y)
    assert h.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
    L.check(L.lib().fg_geglu(h.data_ptr(), h.stride(0), rows, inner, y.data_ptr(), y.stride(0), _stream()), "fg_geglu")


def cast_resample(x: torch.Tensor, x_ld: int, n: int, h: int, w: int, c: int, y: torch.Tensor, y_ld: int,
                  up2: bool = False) -> None:
    """y = bf16(x) for channels-last [n,h,w,c] (fp32 / bf16), optionally nearest-upsampled to [n,2h,2w,c]."""
    _need_cuda(x, y)
    assert y.dtype == torch.bfloat16
    L.check(L.lib().fg_cast_resample(x.data_ptr(), _DT[x.dtype], x_ld, n, h, w, c, L.FG_RES_UP2 if up2 else L.FG_RES_NONE,
                                     y.data_ptr(), y_ld, _stream()), "fg_cast_resample")


def latent_affine(x: torch.Tensor, a: torch.Tensor, b: Optional[torch.Tensor], mode: int, out_dtype: torch.dtype,
                  channel_dim: int = 1) -> torch.Tensor:
    """``fg_latent_affine`` on a contiguous [B, C, ...] latent: mode 0 ``x / a + b``, 1 ``(x - b) * a``, 2 ``x * a``
    (per channel; every op rounded to x's dtype like the reference's eager ops), result in ``out_dtype``."""
    _need_cuda(x, a, b)
    if not x.is_contiguous():
        x = x.contiguous()
    assert a.dtype == torch.float32 and (b is None or b.dtype == torch.float32)
    c = a.numel()
    outer = 1
    for s_ in x.shape[:channel_dim]:
        outer *= s_
    if c == 1:
        outer, inner = 1, x.numel()
    else:
        assert x.shape[channel_dim] == c, (x.shape, c)
        inner = x.numel() // (outer * c)
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    L.check(L.lib().fg_latent_affine(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _DT[out_dtype], a.data_ptr(), _ptr(b),
                                     outer, c, inner, mode, _stream()), "fg_latent_affine")
    return out


def clamp_(x: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    """in-place ``x.clip_(lo, hi)`` as one launch"""
    _need_cuda(x)
    assert x.is_contiguous()
    L.check(L.lib().fg_clamp(x.data_ptr(), _DT[x.dtype], x.numel(), lo, hi, _stream()), "fg_clamp")
    return x


def zero_tail_rows(x: torch.Tensor, lens: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """out[s, t] = x[s, t] for t < lens[s], 0 behind: [B, L, D] -> contiguous [B, L, D] in ``out_dtype``"""
    _need_cuda(x, lens)
    assert x.dim() == 3 and x.stride(2) == 1 and lens.dtype == torch.int64 and lens.numel() == x.shape[0]
    b, rows, d = x.shape
    out = torch.empty(b, rows, d, device=x.device, dtype=out_dtype)
    L.check(L.lib().fg_zero_tail_rows(x.data_ptr(), _DT[x.dtype], x.stride(1), x.stride(0), lens.data_ptr(), b, rows, d,
                                      out.data_ptr(), _DT[out_dtype], _stream()), "fg_zero_tail_rows")
    return out
