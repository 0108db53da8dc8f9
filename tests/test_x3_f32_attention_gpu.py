"""fp32 CUDA-core attention of the split-precision ("bf16x3") EDM path: `fg_attention_f32` against fp32 torch, and the
network with `x3_attention = "f32"` against the reference's golden forwards at the 1e-4 bar BASELINE.json states for the
fp32 path (profiles/x3_emulation.py predicts ~1e-5; with the bf16 tensor-core attention the path measures ~1e-3)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import edm_oracle as EO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _note(line):
    out = os.path.join(os.path.dirname(__file__), "..", "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "x3_f32_attention.txt"), "a") as f:
            f.write(line + "\n")


@pytest.mark.parametrize("n,tokens,heads,d", [(2, 64, 2, 64), (1, 100, 1, 64), (2, 256, 1, 256), (1, 1024, 3, 64),
                                              (1, 70, 2, 128), (1, 192, 1, 192)])
def test_attention_f32_matches_torch(n, tokens, heads, d):
    from fastgen_b200 import ops

    torch.manual_seed(8)
    c = heads * d
    ld = 3 * c + 4
    buf = torch.randn(n, tokens, ld, device="cuda") * 1.5
    qkv = buf[..., :3 * c].double().reshape(n, tokens, 3, heads, d)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    w = torch.einsum("nqhd,nkhd->nhqk", q, k / math.sqrt(d)).softmax(dim=-1)
    ref = torch.einsum("nhqk,nkhd->nqhd", w, v).reshape(n, tokens, c)
    hi = torch.zeros(n, tokens, c + 4, device="cuda", dtype=torch.bfloat16)
    lo = torch.zeros_like(hi)
    ops.attention_f32(buf, ld, n, tokens, heads, d, hi, lo, c + 4)
    torch.cuda.synchronize()
    assert hi[..., c:].abs().max().item() == 0 and lo[..., c:].abs().max().item() == 0
    e_hi, e = _rel(hi[..., :c].float(), ref), _rel(hi[..., :c].float() + lo[..., :c].float(), ref)
    _note(f"kernel n={n} T={tokens} H={heads} d={d}: hi {e_hi:.3e}  hi+lo {e:.3e}")
    assert e_hi < 4e-3 and e < 2e-5, (e_hi, e)
    # hi alone (lo = NULL) is the plain bf16 rounding of the same result
    hi2 = torch.zeros_like(hi)
    ops.attention_f32(buf, ld, n, tokens, heads, d, hi2, None, c + 4)
    assert torch.equal(hi2, hi)


def test_attention_f32_sharp_softmax_and_unsupported_width():
    from fastgen_b200 import _lib as L
    from fastgen_b200 import ops

    torch.manual_seed(9)
    n, tokens, heads, d = 1, 512, 2, 64
    c = heads * d
    x = torch.randn(n, tokens, 3 * c, device="cuda")
    x[:, :, :2 * c] *= 4.0
    x[:, 384:, c:2 * c] *= 2.5   # the running maximum grows in the later key tiles
    qkv = x.double().reshape(n, tokens, 3, heads, d)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    w = torch.einsum("nqhd,nkhd->nhqk", q, k / math.sqrt(d)).softmax(dim=-1)
    ref = torch.einsum("nhqk,nkhd->nqhd", w, v).reshape(n, tokens, c)
    hi = torch.zeros(n, tokens, c, device="cuda", dtype=torch.bfloat16)
    lo = torch.zeros_like(hi)
    ops.attention_f32(x, 3 * c, n, tokens, heads, d, hi, lo, c)
    assert torch.isfinite(hi.float()).all()
    assert _rel(hi.float() + lo.float(), ref) < 5e-5
    with pytest.raises(L.FastGenB200Error, match="not supported"):
        ops.attention_f32(torch.zeros(1, 8, 96, device="cuda"), 96, 1, 8, 1, 32, hi, lo, c)


@pytest.mark.parametrize("name", ["song_tiny", "dhariwal_tiny"])
def test_x3_forward_with_f32_attention_reaches_the_fp32_bar(name):
    from fastgen_b200.edm import EDMPrecond

    g = np.load(os.path.join(GOLD, f"edm_{name}.npz"), allow_pickle=False)
    cfg = json.loads(str(g["cfg"]))
    net = EDMPrecond(compute="bf16x3", **cfg)
    net.load_state_dict(EO.make_state_dict(cfg, int(g["seed"])), strict=False)
    net = net.cuda().eval()
    x_t, t, labels = (torch.from_numpy(np.asarray(g[k])).cuda() for k in ("x_t", "t", "labels"))
    ref = torch.from_numpy(g["fwd_x0"])
    net.set_x3_attention("bf16")
    e_bf16 = _rel(net(x_t, t, condition=labels, fwd_pred_type="x0"), ref)
    net.set_x3_attention("f32")
    out = net(x_t, t, condition=labels, fwd_pred_type="x0")
    e_f32 = _rel(out, ref)
    _note(f"network {name}: x3 + bf16 attention {e_bf16:.3e}   x3 + f32 attention {e_f32:.3e}")
    assert e_f32 < 1e-4, (e_f32, e_bf16)
