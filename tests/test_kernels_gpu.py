"""GPU unit tests of the individual C-ABI kernels against plain PyTorch (fp32/fp64) restatements of the same op.
The end-to-end parity tests against the oracle live in test_edm_parity_gpu.py."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc(x):  # NCHW -> contiguous NHWC
    return x.permute(0, 2, 3, 1).contiguous()


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


CONV_CASES = [
    # n, h, w, cin, cout, taps
    (2, 8, 8, 64, 64, 9),
    (3, 16, 16, 128, 192, 9),
    (2, 32, 32, 192, 384, 9),
    (1, 64, 64, 64, 128, 9),
    (5, 8, 8, 256, 768, 1),
    (2, 4, 4, 32, 48, 9),      # cin < 64: TMA zero-fill of the K chunk
    (2, 16, 16, 8, 64, 9),     # first conv (3 channels padded to 8)
    (2, 16, 16, 192, 3, 9),    # out conv (scalar epilogue, BLOCK_N=16)
    (200, 1, 1, 768, 1000, 1), # linear layer, ragged rows (200 of 256)
    (2, 2, 2, 32, 32, 9),      # the reference tests' tiny resolution
    (1, 128, 128, 64, 64, 9),
]


@pytest.mark.parametrize("n,h,w,cin,cout,taps", CONV_CASES)
def test_conv_gemm_matches_torch(n, h, w, cin, cout, taps):
    from fastgen_b200 import ops

    torch.manual_seed(0)
    dev = "cuda"
    k = 3 if taps == 9 else 1
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, device=dev)
    xb = x.bfloat16()
    wb = wt.bfloat16()
    ref = F.conv2d(xb.float(), wb.float(), bias, padding=k // 2)
    a = _nhwc(xb)
    wk = wb.permute(0, 2, 3, 1).contiguous()  # [cout][kh][kw][cin]
    out = torch.empty(n, h, w, cout, device=dev, dtype=torch.float32)
    ops.conv_gemm([ops.Seg(a, cin, wk, cin, taps)], n, h, w, cout, out, cout, bias=bias)
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2)
    assert _rel(got, ref) < 2e-5, _rel(got, ref)


def test_conv_gemm_fused_tail_and_segments():
    """two K segments (3x3 + 1x1 skip), per-sample vector, fp32 residual, scale, bf16 output, strided views"""
    from fastgen_b200 import ops

    torch.manual_seed(1)
    dev = "cuda"
    n, h, w, c1, c2, cout = 3, 16, 16, 128, 192, 256
    x1 = torch.randn(n, c1, h, w, device=dev).bfloat16()
    x2 = torch.randn(n, c2, h, w, device=dev).bfloat16()
    w1 = (torch.randn(cout, c1, 3, 3, device=dev) / math.sqrt(9 * c1)).bfloat16()
    w2 = (torch.randn(cout, c2, 1, 1, device=dev) / math.sqrt(c2)).bfloat16()
    bias = torch.randn(cout, device=dev)
    rowvec = torch.randn(n, cout + 8, device=dev)
    res = torch.randn(n, h, w, cout + 64, device=dev)
    scale = 0.7071
    ref = (F.conv2d(x1.float(), w1.float(), None, padding=1) + F.conv2d(x2.float(), w2.float()) +
           bias.view(1, -1, 1, 1) + rowvec[:, :cout].reshape(n, cout, 1, 1) + res[..., :cout].permute(0, 3, 1, 2)) * scale
    # operands live inside wider buffers (concat-style strided views)
    buf = torch.zeros(n, h, w, c1 + c2, device=dev, dtype=torch.bfloat16)
    buf[..., :c1] = _nhwc(x1)
    buf[..., c1:] = _nhwc(x2)
    out = torch.zeros(n, h, w, cout + 32, device=dev, dtype=torch.bfloat16)
    ops.conv_gemm(
        [ops.Seg(buf, c1 + c2, w1.permute(0, 2, 3, 1).contiguous(), c1, 9),
         ops.Seg(buf[..., c1:], c1 + c2, w2.permute(0, 2, 3, 1).contiguous(), c2, 1)],
        n, h, w, cout, out, cout + 32, bias=bias, rowvec=rowvec, rowvec_ld=cout + 8,
        res=res, res_ld=cout + 64, res_mode=1, out_scale=scale)
    torch.cuda.synchronize()
    got = out[..., :cout].permute(0, 3, 1, 2).float()
    assert _rel(got, ref) < 4e-3
    assert out[..., cout:].abs().max().item() == 0.0  # nothing written outside the view


@pytest.mark.parametrize("mode", ["up", "down"])
def test_conv_gemm_resampled_residual(mode):
    from fastgen_b200 import ops

    torch.manual_seed(2)
    dev = "cuda"
    n, h, w, c = 2, 16, 16, 64
    x = torch.randn(n, c, h, w, device=dev).bfloat16()
    wt = (torch.randn(c, c, 3, 3, device=dev) / math.sqrt(9 * c)).bfloat16()
    if mode == "up":
        r = torch.randn(n, c, h // 2, w // 2, device=dev)
        rr = F.interpolate(r, scale_factor=2, mode="nearest")
        rm = 2
    else:
        r = torch.randn(n, c, h * 2, w * 2, device=dev)
        rr = F.avg_pool2d(r, 2)
        rm = 3
    ref = F.conv2d(x.float(), wt.float(), padding=1) + rr
    out = torch.empty(n, h, w, c, device=dev)
    ops.conv_gemm([ops.Seg(_nhwc(x), c, wt.permute(0, 2, 3, 1).contiguous(), c, 9)], n, h, w, c, out, c,
                  res=_nhwc(r), res_ld=c, res_mode=rm)
    torch.cuda.synchronize()
    assert _rel(out.permute(0, 3, 1, 2), ref) < 2e-5


RES_RING_CASES = [
    # n, h, w, cin, cout, taps, dtype, mode  (enough tiles that every CTA wraps the slab ring several times)
    (64, 32, 32, 64, 384, 1, torch.float32, "same"),
    (64, 32, 32, 64, 320, 1, torch.float32, "same"),   # ragged last N tile (320 = 192 + 128)
    (40, 16, 16, 128, 576, 9, torch.float32, "same"),
    (64, 32, 32, 64, 384, 1, torch.bfloat16, "same"),
    (48, 32, 32, 64, 384, 1, torch.float32, "up"),
    (300, 8, 8, 64, 768, 1, torch.float32, "up"),      # two images per tile, ragged image count
    (9, 64, 64, 64, 192, 9, torch.float32, "up"),
    (600, 1, 1, 256, 512, 1, torch.float32, "same"),   # linear layer with a residual matrix
]


@pytest.mark.parametrize("n,h,w,cin,cout,taps,dtype,mode", RES_RING_CASES)
def test_conv_gemm_tma_residual_ring(n, h, w, cin, cout, taps, dtype, mode):
    """residual prefetched by TMA into the epilogue slab ring (FG_RES_SAME in place, FG_RES_UP2 from the half-size box)"""
    from fastgen_b200 import ops

    torch.manual_seed(5)
    dev = "cuda"
    k = 3 if taps == 9 else 1
    x = torch.randn(n, cin, h, w, device=dev).bfloat16()
    wt = (torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k)).bfloat16()
    bias = torch.randn(cout, device=dev)
    if mode == "up":
        r = torch.randn(n, cout, h // 2, w // 2, device=dev).to(dtype)
        rr = F.interpolate(r.float(), scale_factor=2, mode="nearest")
        rm = 2
    else:
        r = torch.randn(n, cout, h, w, device=dev).to(dtype)
        rr = r.float()
        rm = 1
    ref = (F.conv2d(x.float(), wt.float(), bias, padding=k // 2) + rr) * 0.5
    out = torch.empty(n, h, w, cout, device=dev, dtype=dtype)
    ops.conv_gemm([ops.Seg(_nhwc(x), cin, wt.permute(0, 2, 3, 1).contiguous(), cin, taps)], n, h, w, cout, out, cout,
                  bias=bias, res=_nhwc(r), res_ld=cout, res_mode=rm, out_scale=0.5)
    torch.cuda.synchronize()
    tol = 2e-5 if dtype == torch.float32 else 4e-3
    assert _rel(out.permute(0, 3, 1, 2), ref) < tol, _rel(out.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("rows,k,n,out_dtype,act,gate_rows,res", [
    (4680, 256, 512, torch.bfloat16, 1, 0, False),      # FFN up: GELU(tanh), bf16 out
    (4680, 512, 256, torch.bfloat16, 0, 1560, True),    # FFN down: per-frame gate * v + residual (bf16 stream)
    (1000, 128, 384, torch.float32, 2, 250, True),      # SiLU + gate + fp32 residual
    (300, 64, 40, torch.float32, 1, 100, True),         # generic (non-TMA) epilogue: cout = 40
])
def test_linear_activation_gate_residual(rows, k, n, out_dtype, act, gate_rows, res):
    """transformer tails: out = gate[row // gate_rows] * act(A.W^T + bias) + residual"""
    from fastgen_b200 import ops

    torch.manual_seed(7)
    dev = "cuda"
    a = torch.randn(rows, k, device=dev).bfloat16()
    wt = (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
    bias = torch.randn(n, device=dev)
    v = a.float() @ wt.float().t() + bias
    if act == 1:
        v = F.gelu(v, approximate="tanh")
    elif act == 2:
        v = F.silu(v)
    kw = {}
    if gate_rows:
        groups = (rows + gate_rows - 1) // gate_rows
        gate = torch.randn(groups, n + 4, device=dev)
        v = v * gate[:, :n].repeat_interleave(gate_rows, 0)[:rows]
        kw.update(gate=gate, gate_ld=n + 4, gate_rows=gate_rows)
    if res:
        r = torch.randn(rows, n, device=dev).to(out_dtype)
        v = v + r.float()
        kw.update(res=r, res_ld=n, res_mode=1)
    out = torch.empty(rows, n, device=dev, dtype=out_dtype)
    ops.conv_gemm([ops.Seg(a, k, wt, k, 1)], rows, 1, 1, n, out, n, bias=bias, act=act, **kw)
    torch.cuda.synchronize()
    tol = 5e-3 if out_dtype == torch.bfloat16 else 1e-3  # tanh.approx in the GELU
    assert _rel(out, v) < tol, _rel(out, v)


@pytest.mark.parametrize("n,h,w,cin,cout,taps,groups,ch0,ctot,res", [
    (20, 32, 32, 64, 192, 9, 32, 0, 192, True),      # cg = 6: groups straddle the 32-column slabs
    (300, 8, 8, 64, 768, 1, 32, 0, 768, False),      # two images per tile, ragged image count
    (6, 64, 64, 32, 384, 9, 32, 0, 384, True),
    (16, 16, 16, 64, 320, 1, 32, 64, 576, False),    # channels [64, 384) of a wider normalised tensor, ragged N tile
])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_conv_gemm_fused_groupnorm_statistics(n, h, w, cin, cout, taps, groups, ch0, ctot, res, out_dtype):
    """the conv epilogue accumulates per (image, group) sum / sum of squares of its output (fp32, or the ROUNDED values of
    a bf16 output -- what the consumer's fg_gn_apply reads -- summed as pairs of adjacent columns)"""
    from fastgen_b200 import ops

    torch.manual_seed(9)
    dev = "cuda"
    k = 3 if taps == 9 else 1
    x = torch.randn(n, cin, h, w, device=dev).bfloat16()
    wt = (torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k)).bfloat16()
    bias = torch.randn(cout, device=dev)
    kw = {}
    if res:
        r = torch.randn(n, h, w, cout, device=dev).to(out_dtype)
        kw = dict(res=r, res_ld=cout, res_mode=1)
    out = torch.empty(n, h, w, cout, device=dev, dtype=out_dtype)
    stats = torch.zeros(n, groups, 2, device=dev, dtype=torch.float64)
    cg = ctot // groups
    ops.conv_gemm([ops.Seg(_nhwc(x), cin, wt.permute(0, 2, 3, 1).contiguous(), cin, taps)], n, h, w, cout, out, cout,
                  bias=bias, out_scale=0.7, stats=stats, stats_groups=groups, stats_cg=cg, stats_ch0=ch0, **kw)
    torch.cuda.synchronize()
    ref = torch.zeros(n, groups, 2, device=dev, dtype=torch.float64)
    o = out.double()
    gidx = (ch0 + torch.arange(cout, device=dev)) // cg
    ref[:, :, 0].index_add_(1, gidx, o.sum(dim=(1, 2)))
    ref[:, :, 1].index_add_(1, gidx, (o * o).sum(dim=(1, 2)))
    assert torch.allclose(stats, ref, rtol=2e-5, atol=1e-3), (stats - ref).abs().max().item()


def test_conv_gemm_two_statistics_targets():
    """a tensor that is also a skip connection: the epilogue feeds its own next norm (stats) and, at a channel offset with
    another group width and group count, the decoder norm that reads it inside a concatenation (stats2)"""
    from fastgen_b200 import ops

    torch.manual_seed(11)
    dev = "cuda"
    n, h, w, cin, cout = 5, 16, 16, 64, 192
    x = torch.randn(n, cin, h, w, device=dev).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)).bfloat16()
    out = torch.empty(n, h, w, cout, device=dev)
    g1, cg1 = 32, cout // 32                      # own norm: 32 groups of 6
    g2, ctot, ch0 = 16, 576, 384                  # concat norm: 16 groups of 36 over 576 channels, ours are [384, 576)
    cg2 = ctot // g2
    s1 = torch.zeros(n, g1, 2, device=dev, dtype=torch.float64)
    s2 = torch.full((n, g2, 2), 3.0, device=dev, dtype=torch.float64)   # the other producer has already added its share
    ops.conv_gemm([ops.Seg(_nhwc(x), cin, wt.permute(0, 2, 3, 1).contiguous(), cin, 9)], n, h, w, cout, out, cout,
                  stats=s1, stats_groups=g1, stats_cg=cg1, stats_ch0=0,
                  stats2=s2, stats2_groups=g2, stats2_cg=cg2, stats2_ch0=ch0)
    torch.cuda.synchronize()
    o = out.double()
    for st, groups, cg, c0, base in ((s1, g1, cg1, 0, 0.0), (s2, g2, cg2, ch0, 3.0)):
        ref = torch.full((n, groups, 2), base, device=dev, dtype=torch.float64)
        gidx = (c0 + torch.arange(cout, device=dev)) // cg
        ref[:, :, 0].index_add_(1, gidx, o.sum(dim=(1, 2)))
        ref[:, :, 1].index_add_(1, gidx, (o * o).sum(dim=(1, 2)))
        assert torch.allclose(st, ref, rtol=2e-5, atol=1e-3), (st - ref).abs().max().item()


def test_conv_gemm_split_precision_segments():
    """bf16x3 split products (hi*hi + lo*hi + hi*lo) reach ~fp32 accuracy: the 'fp32 path' of the network"""
    from fastgen_b200 import ops

    torch.manual_seed(3)
    dev = "cuda"
    n, h, w, cin, cout = 2, 16, 16, 128, 128
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(9 * cin)
    ref = F.conv2d(x.double(), wt.double(), padding=1)
    a = _nhwc(x)
    wk = wt.permute(0, 2, 3, 1).contiguous()
    a_hi = a.bfloat16()
    a_lo = (a - a_hi.float()).bfloat16()
    w_hi = wk.bfloat16()
    w_lo = (wk - w_hi.float()).bfloat16()
    out = torch.empty(n, h, w, cout, device=dev)
    ops.conv_gemm([ops.Seg(a_hi, cin, w_hi, cin, 9), ops.Seg(a_lo, cin, w_hi, cin, 9), ops.Seg(a_hi, cin, w_lo, cin, 9)],
                  n, h, w, cout, out, cout)
    torch.cuda.synchronize()
    assert _rel(out.permute(0, 3, 1, 2), ref) < 3e-5


def test_conv_gemm_rejects_bad_arguments():
    from fastgen_b200 import ops

    dev = "cuda"
    a = torch.zeros(1, 6, 6, 64, device=dev, dtype=torch.bfloat16)
    wk = torch.zeros(64, 3, 3, 64, device=dev, dtype=torch.bfloat16)
    out = torch.zeros(1, 6, 6, 64, device=dev)
    with pytest.raises(ValueError):
        ops.conv_gemm([ops.Seg(a, 64, wk, 64, 9)], 1, 6, 6, 64, out, 64)  # w=6 unsupported
    with pytest.raises(ValueError):
        ops.conv_gemm([ops.Seg(a, 64, wk, 64, 4)], 1, 8, 8, 64, out, 64)  # taps


GN_CASES = [
    # n, h, w, c, groups, dtype, adaptive, silu, resample
    (3, 16, 16, 192, 32, torch.float32, True, True, 0),
    (2, 8, 8, 768, 32, torch.float32, False, True, 0),
    (2, 32, 32, 576, 32, torch.bfloat16, True, True, 0),   # 18 channels/group: vectors straddle groups
    (2, 16, 16, 64, 16, torch.float32, False, False, 0),
    (2, 8, 8, 128, 32, torch.float32, True, True, 2),      # up2
    (2, 16, 16, 128, 32, torch.float32, False, True, 3),   # down2
    (1, 64, 64, 1344, 32, torch.float32, True, True, 0),
]


@pytest.mark.parametrize("n,h,w,c,groups,dtype,adaptive,silu,resample", GN_CASES)
def test_groupnorm_matches_torch(n, h, w, c, groups, dtype, adaptive, silu, resample):
    from fastgen_b200 import ops

    torch.manual_seed(4)
    dev = "cuda"
    ld = c + 16
    xbuf = torch.randn(n, h, w, ld, device=dev) * 1.7 + 0.6
    xbuf = xbuf.to(dtype)
    x = xbuf[..., :c]
    gamma = torch.randn(c, device=dev)
    beta = torch.randn(c, device=dev)
    eps = 1e-5
    xn = x.float().permute(0, 3, 1, 2)
    ref = F.group_norm(xn, groups, gamma, beta, eps)
    scale = shift = None
    if adaptive:
        mod = torch.randn(n, 2 * c + 8, device=dev)
        scale, shift = mod[:, :c], mod[:, c:2 * c]
        ref = torch.addcmul(shift.reshape(n, c, 1, 1), ref, scale.reshape(n, c, 1, 1) + 1)
    if silu:
        ref = F.silu(ref)
    raw_ref = xn
    oh, ow = h, w
    if resample == 2:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
        raw_ref = F.interpolate(raw_ref, scale_factor=2, mode="nearest")
        oh, ow = 2 * h, 2 * w
    elif resample == 3:
        ref = F.avg_pool2d(ref, 2)
        raw_ref = F.avg_pool2d(raw_ref, 2)
        oh, ow = h // 2, w // 2
    stats = torch.empty(n, groups, 2, device=dev, dtype=torch.float64)
    ops.gn_stats(xbuf, ld, n, h * w, c, groups, stats)
    y = torch.zeros(n, oh, ow, c, device=dev, dtype=torch.bfloat16)
    y_lo = torch.zeros_like(y)
    raw = torch.zeros_like(y)
    raw_lo = torch.zeros_like(y)
    ops.gn_apply(xbuf, ld, n, h, w, c, groups, stats, eps, gamma, beta, y, c,
                 scale=scale, shift=shift if adaptive else None, mod_ld=(2 * c + 8) if adaptive else 0,
                 silu=silu, resample=resample, y_lo=y_lo, raw=raw, raw_ld=c, raw_lo=raw_lo)
    torch.cuda.synchronize()
    # shift view starts at column c of `mod`: pass the base pointer of the view
    got = (y.float() + y_lo.float()).permute(0, 3, 1, 2)
    got_hi = y.float().permute(0, 3, 1, 2)
    assert _rel(got_hi, ref) < 4e-3
    assert _rel(got, ref) < 3e-5
    assert _rel((raw.float() + raw_lo.float()).permute(0, 3, 1, 2), raw_ref) < 2e-5
    # statistics themselves
    cnt = h * w * (c // groups)
    xg = x.double().reshape(n, h * w, groups, c // groups)
    assert torch.allclose(stats[..., 0], xg.sum(dim=(1, 3)), rtol=1e-5, atol=1e-3 * math.sqrt(cnt))
    assert torch.allclose(stats[..., 1], (xg * xg).sum(dim=(1, 3)), rtol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["last", "sde", "ode"])
def test_edm_step_bit_exact(dtype, mode):
    """precond_output + re-noise is bit-identical to the reference formulas evaluated with torch on the GPU"""
    from fastgen_b200 import ops

    torch.manual_seed(5)
    dev = "cuda"
    n, c, h, w = 5, 3, 16, 16
    sd = 0.5
    x_t = (torch.randn(n, c, h, w, device=dev) * 3).to(dtype)
    eps = torch.randn(n, c, h, w, device=dev).to(dtype)
    f = torch.randn(n, h, w, 4, device=dev)
    t = torch.tensor([80.0, 2.5265, 0.17, 1.1, 17.498], device=dev, dtype=torch.float64)
    t_next = torch.tensor([1.1, 0.5, 0.002, 0.3, 2.0], device=dev, dtype=torch.float64)
    c_skip = sd**2 / (t**2 + sd**2)
    c_out = t * sd / (t**2 + sd**2).sqrt()
    F_ = f[..., :c].permute(0, 3, 1, 2).to(dtype)
    bc = lambda v: v.reshape(n, 1, 1, 1)
    x0_ref = bc(c_skip.to(dtype)) * x_t + bc(c_out.to(dtype)) * F_
    x0 = torch.empty_like(x_t)
    ones = torch.ones_like(t)
    if mode == "last":
        ops.edm_step(f, 4, x_t, x0, c_skip, c_out)
        torch.cuda.synchronize()
        assert torch.equal(x0, x0_ref)
        return
    x_next = torch.empty_like(x_t)
    if mode == "sde":
        e = eps
        ops.edm_step(f, 4, x_t, x0, c_skip, c_out, x_next=x_next, eps=eps, alpha_next=ones, sigma_next=t_next)
    else:
        e = ((x_t.double() - x0_ref.double() * bc(ones)) / bc(t)).to(dtype)
        ops.edm_step(f, 4, x_t, x0, c_skip, c_out, x_next=x_next, alpha_next=ones, sigma_next=t_next, ode=True,
                     alpha_cur=ones, sigma_cur=t)
    torch.cuda.synchronize()
    xn_ref = (x0_ref.double() * bc(ones) + e.double() * bc(t_next)).to(dtype)
    assert torch.equal(x0, x0_ref)
    assert torch.equal(x_next, xn_ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
def test_axpby_and_sub_div_bit_exact(dtype):
    from fastgen_b200 import ops

    torch.manual_seed(6)
    dev = "cuda"
    a = torch.randn(4, 16, 3, 10, 12, device=dev).to(dtype)
    b = torch.randn_like(a)
    ca = torch.rand(4, device=dev, dtype=torch.float64)
    cb = torch.rand(4, device=dev, dtype=torch.float64) + 0.1
    bc = lambda v: v.reshape(4, 1, 1, 1, 1)
    assert torch.equal(ops.axpby(a, ca, b, cb), (a.double() * bc(ca) + b.double() * bc(cb)).to(dtype))
    assert torch.equal(ops.axpby(a, ca), (a.double() * bc(ca)).to(dtype))
    assert torch.equal(ops.sub_div(a, b, ca, cb), ((a.double() - b.double() * bc(ca)) / bc(cb)).to(dtype))


def test_edm_input_layout_and_scale():
    from fastgen_b200 import ops

    torch.manual_seed(7)
    dev = "cuda"
    n, c, h, w = 3, 3, 8, 8
    x = torch.randn(n, c, h, w, device=dev) * 40
    t = torch.tensor([80.0, 1.1, 0.3], device=dev, dtype=torch.float64)
    c_in = 1 / (0.25 + t**2).sqrt()
    xin = torch.empty(n, h, w, 8, device=dev, dtype=torch.bfloat16)
    lo = torch.empty_like(xin)
    ops.edm_input(x, c_in, xin, 8, xin_lo=lo)
    torch.cuda.synchronize()
    ref = (c_in.float().reshape(n, 1, 1, 1) * x).permute(0, 2, 3, 1)
    assert torch.equal(xin[..., :3], ref.bfloat16())
    assert xin[..., 3:].abs().max().item() == 0
    assert _rel(xin[..., :3].float() + lo[..., :3].float(), ref) < 1e-5


def test_pos_embedding_and_split():
    from fastgen_b200 import ops

    dev = "cuda"
    t = torch.tensor([1.0955, -1.2, 0.02], device=dev)
    for ch, endpoint, swap in [(128, True, True), (192, False, False)]:
        out = torch.empty(3, ch, device=dev)
        ops.pos_embedding(t, ch, 10000, endpoint, swap, out, ch)
        freqs = torch.arange(ch // 2, dtype=torch.float32, device=dev) / (ch // 2 - (1 if endpoint else 0))
        freqs = (1 / 10000) ** freqs
        e = t.ger(freqs)
        ref = torch.cat([e.cos(), e.sin()], 1)
        if swap:
            ref = ref.reshape(3, 2, -1).flip(1).reshape(3, ch)
        assert torch.allclose(out, ref, atol=2e-6)
    x = torch.randn(5, 40, device=dev)
    add = torch.randn(5, 40, device=dev)
    hi = torch.empty(5, 40, device=dev, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    y = torch.empty(5, 40, device=dev)
    ops.split_bf16(x, 40, 5, 40, hi, 40, lo=lo, add=add, add_ld=40, act=1, y=y, y_ld=40)
    ref = F.silu(x + add)
    assert torch.allclose(y, ref, atol=1e-6)
    assert _rel(hi.float() + lo.float(), ref) < 1e-5


ATTN_CASES = [
    # n, tokens, heads, d
    (2, 1024, 6, 64),
    (3, 256, 9, 64),
    (2, 64, 12, 64),
    (2, 256, 1, 256),
    (2, 64, 1, 256),
    (1, 16, 2, 32),     # tiny nets: d < 64 (TMA zero fill), T < one tile
    (2, 4, 1, 64),
    (1, 384, 2, 128),
    (1, 4096, 2, 64),   # SDXL-like token count
]


@pytest.mark.parametrize("n,tokens,heads,d", ATTN_CASES)
def test_attention_matches_torch(n, tokens, heads, d):
    from fastgen_b200 import ops

    torch.manual_seed(8)
    dev = "cuda"
    c = heads * d
    ld = 3 * c + 8
    buf = (torch.randn(n, tokens, ld, device=dev) * 1.5).bfloat16()
    qkv = buf[..., :3 * c].float().reshape(n, tokens, 3, heads, d)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]  # [n, T, H, d]
    w = torch.einsum("nqhd,nkhd->nhqk", q, k / math.sqrt(d)).softmax(dim=-1)
    ref = torch.einsum("nhqk,nkhd->nqhd", w, v).reshape(n, tokens, c)
    out = torch.zeros(n, tokens, c + 8, device=dev, dtype=torch.bfloat16)
    ops.attention(buf, ld, n, tokens, heads, d, out, c + 8)
    torch.cuda.synchronize()
    assert out[..., c:].abs().max().item() == 0
    assert _rel(out[..., :c], ref) < 1e-2, _rel(out[..., :c], ref)


def test_attention_large_logits_rescale():
    """running-max growth across KV tiles exercises the lazy O rescale"""
    from fastgen_b200 import ops

    torch.manual_seed(9)
    dev = "cuda"
    n, tokens, heads, d = 1, 512, 2, 64
    c = heads * d
    x = torch.randn(n, tokens, 3 * c, device=dev)
    x[:, :, :2 * c] *= 4.0                      # sharp softmax
    x[:, 384:, c:2 * c] *= 2.5                  # later keys carry much larger logits
    buf = x.bfloat16()
    qkv = buf.float().reshape(n, tokens, 3, heads, d)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    w = torch.einsum("nqhd,nkhd->nhqk", q, k / math.sqrt(d)).softmax(dim=-1)
    ref = torch.einsum("nhqk,nkhd->nqhd", w, v).reshape(n, tokens, c)
    out = torch.zeros(n, tokens, c, device=dev, dtype=torch.bfloat16)
    ops.attention(buf, 3 * c, n, tokens, heads, d, out, c)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 1.5e-2, _rel(out, ref)


@pytest.mark.parametrize("rows,k,inner", [(300, 64, 128), (1000, 640, 2560), (77, 128, 200)])
def test_conv_gemm_geglu_tail(rows, k, inner):
    """FG_ACT_GEGLU: (hidden, gate) column pairs -> hidden * gelu_erf(gate), half-width bf16 output (SDXL ff.net.0)"""
    import torch.nn.functional as F

    from fastgen_b200 import _lib as L
    from fastgen_b200 import ops

    torch.manual_seed(0)
    a = torch.randn(rows, k, device="cuda").bfloat16()
    w = (torch.randn(2 * inner, k, device="cuda") * k ** -0.5).bfloat16()       # rows [0, inner) hidden, then gate
    bias = 0.1 * torch.randn(2 * inner, device="cuda")
    perm = torch.stack([torch.arange(inner), torch.arange(inner) + inner], 1).reshape(-1).cuda()
    out = torch.zeros(rows, inner + 8, device="cuda", dtype=torch.bfloat16)       # ld > inner
    ops.conv_gemm([ops.Seg(a, k, w[perm].contiguous(), k, 1)], rows, 1, 1, 2 * inner, out, inner + 8,
                  bias=bias[perm].contiguous(), act=L.FG_ACT_GEGLU)
    torch.cuda.synchronize()
    v = a.float() @ w.float().t() + bias
    ref = v[:, :inner] * F.gelu(v[:, inner:])
    got = out[:, :inner].float()
    assert ((got - ref).norm() / ref.norm()).item() < 4e-3
    assert torch.all(out[:, inner:] == 0)
