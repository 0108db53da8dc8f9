"""GPU unit tests of the transformer (Wan / DiT) C-ABI kernels against plain PyTorch restatements of the same op."""

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows,d,mode,rpg", [
    (4680, 1536, "mod", 1560), (4680, 1536, "affine", 0), (1000, 1152, "mod", 1000), (333, 5120, "mod", 111),
    (64, 64, "none", 0),
])
def test_ln_modulate(rows, d, mode, rpg):
    from fastgen_b200 import ops

    torch.manual_seed(0)
    x = (torch.randn(rows, d, device=DEV) * 3 + 0.5).bfloat16()
    y = torch.empty_like(x)
    kw = {}
    ref = F.layer_norm(x.float(), (d,), eps=1e-6)
    if mode == "mod":
        groups = (rows + rpg - 1) // rpg
        mod = torch.randn(groups, 2 * d + 8, device=DEV)
        scale, shift = mod[:, :d], mod[:, d:2 * d]
        ref = ref * (1 + scale.repeat_interleave(rpg, 0)[:rows]) + shift.repeat_interleave(rpg, 0)[:rows]
        kw = dict(scale=scale, shift=shift, mod_ld=2 * d + 8, rows_per_group=rpg)
    elif mode == "affine":
        g, b = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
        ref = ref * g + b
        kw = dict(gamma=g, beta=b)
    ops.ln_modulate(x, d, rows, d, 1e-6, y, d, **kw)
    torch.cuda.synchronize()
    assert _rel(y, ref.bfloat16()) < 3e-3, _rel(y, ref.bfloat16())


def _rope_tables(d, max_pos=1024, theta=10000.0):
    """WanRotaryPosEmbed tables reduced to one entry per pair: [max_pos, d/2] (cos[..., 0::2] / sin[..., 1::2])."""
    h_dim = w_dim = 2 * (d // 6)
    t_dim = d - h_dim - w_dim
    cos, sin = [], []
    pos = torch.arange(max_pos, dtype=torch.float64)
    for dim in (t_dim, h_dim, w_dim):
        fr = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
        ang = torch.outer(pos, fr)
        cos.append(ang.cos().float())
        sin.append(ang.sin().float())
    return torch.cat(cos, 1).contiguous(), torch.cat(sin, 1).contiguous(), t_dim // 2, h_dim // 2


def _rope_ref(x, cos, sin, nt, nh, grid, frame0):
    """x [B, N, H, d] bf16 -> roped, reference formulation (network_causal.py:274-286)."""
    f, hh, ww = grid
    idx = torch.arange(f * hh * ww, device=x.device)
    pf, ph, pw = frame0 + idx // (hh * ww), (idx // ww) % hh, idx % ww
    c = torch.cat([cos[pf, :nt], cos[ph, nt:nt + nh], cos[pw, nt + nh:]], 1)[None, :, None, :]
    s = torch.cat([sin[pf, :nt], sin[ph, nt:nt + nh], sin[pw, nt + nh:]], 1)[None, :, None, :]
    x1, x2 = x.unflatten(-1, (-1, 2)).unbind(-1)
    out = torch.empty_like(x)
    out[..., 0::2] = x1 * c - x2 * s
    out[..., 1::2] = x1 * s + x2 * c
    return out


@pytest.mark.parametrize("b,grid,heads,d,frame0,cache", [(2, (3, 6, 10), 12, 128, 4, True), (1, (2, 4, 4), 4, 64, 0, False),
                                                         (1, (3, 30, 52), 12, 128, 18, True)])
def test_qk_norm_rope_and_cache_append(b, grid, heads, d, frame0, cache):
    from fastgen_b200 import ops

    torch.manual_seed(1)
    n = grid[0] * grid[1] * grid[2]
    dm = heads * d
    qkv = torch.randn(b * n, 3 * dm, device=DEV).bfloat16()
    wq = (1 + 0.1 * torch.randn(dm, device=DEV)).bfloat16()
    wk = (1 + 0.1 * torch.randn(dm, device=DEV)).bfloat16()
    cos, sin, nt, nh = [t.to(DEV) if isinstance(t, torch.Tensor) else t for t in _rope_tables(d)]
    rms = torch.nn.RMSNorm(dm, eps=1e-6, device=DEV, dtype=torch.bfloat16)

    def ref(x, w):
        rms.weight.data.copy_(w)
        y = rms(x).view(b, n, heads, d)
        return _rope_ref(y, cos, sin, nt, nh, grid, frame0).reshape(b * n, dm)

    q_ref, k_ref = ref(qkv[:, :dm], wq), ref(qkv[:, dm:2 * dm], wk)
    v_ref = qkv[:, 2 * dm:].clone()
    cap = frame0 * grid[1] * grid[2] + n + 7
    kc = torch.zeros(b, cap, dm, device=DEV, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    row0 = frame0 * grid[1] * grid[2]
    kw = dict(v=qkv[:, 2 * dm:], v_ld=3 * dm, k_cache=kc, v_cache=vc, cache_ld=dm, cache_bs=cap * dm,
              cache_row0=row0) if cache else {}
    ops.qk_norm_rope(qkv, qkv[:, dm:], 3 * dm, 3 * dm, b * n, heads, d, 1e-6, wq, wk, n, rope=(cos, sin, nt, nh),
                     grid=grid, frame0=frame0, **kw)
    torch.cuda.synchronize()
    assert _rel(qkv[:, :dm], q_ref) < 4e-3, _rel(qkv[:, :dm], q_ref)
    assert _rel(qkv[:, dm:2 * dm], k_ref) < 4e-3
    assert torch.equal(qkv[:, 2 * dm:], v_ref)
    if cache:
        assert torch.equal(kc[:, row0:row0 + n].reshape(b * n, dm), qkv[:, dm:2 * dm])
        assert torch.equal(vc[:, row0:row0 + n].reshape(b * n, dm), v_ref)
        assert kc[:, :row0].abs().max().item() == 0 and kc[:, row0 + n:].abs().max().item() == 0


def test_qk_norm_only_for_cross_attention():
    from fastgen_b200 import ops

    torch.manual_seed(2)
    rows, heads, d = 512, 12, 128
    dm = heads * d
    k = torch.randn(rows, dm, device=DEV).bfloat16()
    w = (1 + 0.1 * torch.randn(dm, device=DEV)).bfloat16()
    rms = torch.nn.RMSNorm(dm, eps=1e-6, device=DEV, dtype=torch.bfloat16)
    rms.weight.data.copy_(w)
    ref = rms(k)
    ops.qk_norm_rope(None, k, 0, dm, rows, heads, d, 1e-6, None, w, rows)
    torch.cuda.synchronize()
    assert _rel(k, ref) < 4e-3


def _sdpa(q, k, v, mask=None):
    # q [B,Tq,H,d], k/v [B,Tk,H,d] -> [B,Tq,H*d], fp32 math on the bf16 values
    o = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2),
                                       attn_mask=mask)
    return o.transpose(1, 2).flatten(2)


@pytest.mark.parametrize("b,tq,heads,d,l0,l1", [
    (2, 300, 4, 128, 300, 0),          # plain self-attention, ragged tiles
    (1, 468, 12, 128, 1000, 468),      # KV-cache prefix (ragged, 1000 % 128 != 0) + current chunk
    (2, 200, 3, 64, 130, 200),
    (1, 700, 2, 128, 512, 0),          # cross-attention shape: 512 keys
    (1, 130, 2, 72, 77, 0),            # head dim padded to 128 inside the kernel
])
def test_flash_attn_segments(b, tq, heads, d, l0, l1):
    from fastgen_b200 import ops

    torch.manual_seed(3)
    dm = heads * d
    q = torch.randn(b, tq, dm, device=DEV).bfloat16()
    cap = l0 + 50
    kc = torch.randn(b, cap, dm, device=DEV).bfloat16()  # cache with spare capacity (strides differ from q)
    vc = torch.randn(b, cap, dm, device=DEV).bfloat16()
    segs = [(kc, vc, dm, cap * dm, l0)]
    k_all, v_all = kc[:, :l0], vc[:, :l0]
    if l1:
        kv1 = torch.randn(b, l1, 2 * dm, device=DEV).bfloat16()
        segs.append((kv1, kv1[..., dm:], 2 * dm, l1 * 2 * dm, l1))
        k_all = torch.cat([k_all, kv1[..., :dm]], 1)
        v_all = torch.cat([v_all, kv1[..., dm:]], 1)
    out = torch.empty(b, tq, dm, device=DEV, dtype=torch.bfloat16)
    ops.flash_attn(q, dm, tq * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    torch.cuda.synchronize()
    ref = _sdpa(q.view(b, tq, heads, d), k_all.reshape(b, -1, heads, d), v_all.reshape(b, -1, heads, d))
    assert _rel(out, ref) < 1e-2, _rel(out, ref)
    # with caller-provided scratch the library may split the key range over several CTAs and merge the partials
    p = ops.flash_attn_params(q, dm, tq * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    wsz = ops.flash_attn_workspace_size(p)
    if tq > 128 and d <= 128:
        assert wsz > 0
        ws = torch.empty(wsz, device=DEV, dtype=torch.uint8)
        out2 = torch.zeros_like(out)
        ops.flash_attn(q, dm, tq * dm, segs, b, heads, d, tq, out2, dm, tq * dm, workspace=ws)
        torch.cuda.synchronize()
        assert _rel(out2, ref) < 1e-2, _rel(out2, ref)


@pytest.mark.parametrize("frames,chunk,fs", [(7, 3, 40), (6, 3, 130), (5, 2, 97)])
def test_flash_attn_block_causal_mask(frames, chunk, fs):
    """mask(q, kv) = kv < end_of_chunk(q): first chunk takes the remainder frames (network_causal.py:131-196)"""
    from fastgen_b200 import ops

    torch.manual_seed(4)
    b, heads, d = 1, 2, 128
    n = frames * fs
    dm = heads * d
    qkv = torch.randn(b, n, 3 * dm, device=DEV).bfloat16()
    first = (chunk + frames % chunk) * fs if frames // chunk else (frames % chunk) * fs
    ends = torch.zeros(n, dtype=torch.long, device=DEV)
    cur, counts = 0, ([chunk + frames % chunk] + [chunk] * (frames // chunk - 1))
    for c in counts:
        ends[cur:cur + c * fs] = cur + c * fs
        cur += c * fs
    idx = torch.arange(n, device=DEV)
    mask = idx[None, :] < ends[:, None]
    out = torch.empty(b, n, dm, device=DEV, dtype=torch.bfloat16)
    ops.flash_attn(qkv, 3 * dm, n * 3 * dm, [(qkv[..., dm:], qkv[..., 2 * dm:], 3 * dm, n * 3 * dm, n)], b, heads, d, n,
                   out, dm, n * dm, mask_first=first, mask_chunk=chunk * fs)
    torch.cuda.synchronize()
    ref = _sdpa(qkv[..., :dm].reshape(b, n, heads, d), qkv[..., dm:2 * dm].reshape(b, n, heads, d),
                qkv[..., 2 * dm:].reshape(b, n, heads, d), mask=mask)
    assert _rel(out, ref) < 1e-2, _rel(out, ref)


def test_patchify_unpatchify_roundtrip_and_conv3d():
    from fastgen_b200 import ops

    torch.manual_seed(5)
    b, c, f, h, w, dmod = 2, 16, 3, 12, 20, 128
    x = torch.randn(b, c, f, h, w, device=DEV)
    conv = torch.nn.Conv3d(c, dmod, (1, 2, 2), (1, 2, 2)).to(DEV)
    ref = conv(x.bfloat16().float()).flatten(2).transpose(1, 2)  # [B, N, D]
    rows = b * f * (h // 2) * (w // 2)
    a = torch.empty(rows, c * 4, device=DEV, dtype=torch.bfloat16)
    ops.patchify(x, 2, 2, a, c * 4)
    wt = conv.weight.detach().reshape(dmod, c * 4).bfloat16().contiguous()
    out = torch.empty(rows, dmod, device=DEV)
    ops.conv_gemm([ops.Seg(a, c * 4, wt, c * 4, 1)], rows, 1, 1, dmod, out, dmod, bias=conv.bias.detach().float())
    torch.cuda.synchronize()
    ref2 = F.conv3d(x.bfloat16().float(), wt.float().view(dmod, c, 1, 2, 2), conv.bias, stride=(1, 2, 2)).flatten(2).transpose(1, 2)
    assert _rel(out.view(b, -1, dmod), ref2) < 2e-5
    assert _rel(out.view(b, -1, dmod), ref) < 1e-2
    # unpatchify follows the reference permute (0,7,1,4,2,5,3,6)
    y = torch.randn(rows, 4 * c, device=DEV).bfloat16()
    hs = y.view(b, f, h // 2, w // 2, 1, 2, 2, c).permute(0, 7, 1, 4, 2, 5, 3, 6).flatten(6, 7).flatten(4, 5).flatten(2, 3)
    for dt in (torch.bfloat16, torch.float32):
        o = torch.empty(b, c, f, h, w, device=DEV, dtype=dt)
        ops.unpatchify(y, 4 * c, o, 2, 2)
        torch.cuda.synchronize()
        assert torch.equal(o, hs.to(dt))


def test_bcast_add():
    from fastgen_b200 import ops

    a, b = torch.randn(30, 96, device=DEV), torch.randn(7, 96, device=DEV)
    out = torch.empty(30, 7, 96, device=DEV)
    ops.bcast_add(a, b, out)
    torch.cuda.synchronize()
    assert torch.equal(out, a[:, None] + b[None])


@pytest.mark.parametrize("rows,inner", [(300, 256), (77, 40), (4096, 2560)])
def test_geglu(rows, inner):
    """diffusers GEGLU (SDXL feed-forward): hidden * gelu(gate), gelu evaluated and rounded in bf16"""
    from fastgen_b200 import ops

    torch.manual_seed(0)
    h = (2 * torch.randn(rows, 2 * inner + 8, device=DEV)).bfloat16()   # ld > 2 * inner
    y = torch.zeros(rows, inner, device=DEV, dtype=torch.bfloat16)
    ops.geglu(h, rows, inner, y)
    torch.cuda.synchronize()
    hid, gate = h[:, :inner], h[:, inner:2 * inner]
    ref = hid.float() * F.gelu(gate.float()).bfloat16().float()
    assert _rel(y, ref) < 4e-3, _rel(y, ref)
    assert torch.equal(y, ref.bfloat16()) or (y.float() - ref).abs().max() <= 2 ** -7 * ref.abs().max()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("up2", [False, True])
def test_cast_resample(dt, up2):
    from fastgen_b200 import ops

    torch.manual_seed(1)
    n, h, w, c, ld = 3, 8, 16, 40, 56
    buf = torch.randn(n, h, w, ld, device=DEV).to(dt)
    x = buf[..., 8:8 + c]                                             # a slice of a wider concat buffer
    s = 2 if up2 else 1
    y = torch.zeros(n, s * h, s * w, c, device=DEV, dtype=torch.bfloat16)
    ops.cast_resample(x, ld, n, h, w, c, y, c, up2=up2)
    torch.cuda.synchronize()
    ref = x.bfloat16()
    if up2:
        ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(y, ref)


@pytest.mark.parametrize("rows,d", [(100, 640), (37, 1280), (9, 72)])
def test_ln_affine_fp32_rows(rows, d):
    """fp32 token stream of the SDXL transformer blocks: LayerNorm(x) * gamma + beta, fp32 in, bf16 out"""
    from fastgen_b200 import ops

    torch.manual_seed(2)
    x = 3 * torch.randn(rows, d + 8, device=DEV) + 1.0
    gamma, beta = 1 + 0.1 * torch.randn(d, device=DEV), 0.1 * torch.randn(d, device=DEV)
    y = torch.zeros(rows, d, device=DEV, dtype=torch.bfloat16)
    p = ops.ln_mod_params(x, d + 8, rows, d, 1e-5, y, d, gamma=gamma, beta=beta)
    from fastgen_b200 import _lib as L
    import ctypes
    L.check(L.lib().fg_ln_modulate(ctypes.byref(p), torch.cuda.current_stream().cuda_stream), "fg_ln_modulate")
    torch.cuda.synchronize()
    ref = F.layer_norm(x[:, :d], (d,), gamma, beta, 1e-5)
    assert _rel(y, ref) < 3e-3, _rel(y, ref)


@pytest.mark.parametrize("tq,heads,d", [(300, 8, 40), (64, 8, 160), (1024, 8, 80)])
def test_flash_attn_per_sample_key_lengths(tq, heads, d):
    """a prefix `encoder_attention_mask` (SD 1.5: 77 text tokens, padding masked, fastgen/networks/SD15/network.py:142-145)
    as per-sample key counts: equal to SDPA with the reference's additive -10000 bias on the masked keys"""
    from fastgen_b200 import ops

    torch.manual_seed(7)
    b, lk = 3, 77
    dm = heads * d
    q = torch.randn(b, tq, dm, device=DEV).bfloat16()
    kv = torch.randn(b, lk, 2 * dm, device=DEV).bfloat16()
    lens = torch.tensor([77, 9, 31], dtype=torch.int32, device=DEV)
    out = torch.empty(b, tq, dm, device=DEV, dtype=torch.bfloat16)
    ops.flash_attn(q, dm, tq * dm, [(kv, kv[..., dm:], 2 * dm, lk * 2 * dm, lk)], b, heads, d, tq, out, dm, tq * dm,
                   kv_len0_per_sample=lens)
    torch.cuda.synchronize()
    mask = (torch.arange(lk, device=DEV)[None, :] < lens[:, None]).float()          # [b, lk] of 0 / 1
    bias = ((1 - mask) * -10000.0)[:, None, None, :]                                   # the reference's conversion
    ref = F.scaled_dot_product_attention(q.float().view(b, tq, heads, d).transpose(1, 2),
                                         kv[..., :dm].float().reshape(b, lk, heads, d).transpose(1, 2),
                                         kv[..., dm:].float().reshape(b, lk, heads, d).transpose(1, 2), attn_mask=bias)
    assert _rel(out, ref.transpose(1, 2).flatten(2)) < 1e-2
