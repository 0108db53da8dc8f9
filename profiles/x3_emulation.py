"""CPU emulation of the rounding points of the split-precision ("bf16x3", fp32-path) EDM forward.

Question it answers without a GPU: which rounding points of `EDMPrecond(compute="bf16x3")` keep it from the 1e-4 relative
L2 bar BASELINE.json states for the fp32 path (measured on B200: <= 2e-3), and does fp32 attention close the gap?

The fp32 oracle (oracle/edm_oracle.py) is re-run with its three contraction sites replaced by emulations of what the CUDA
path does:
  * conv / linear: operands split into bf16 hi + lo (16 mantissa bits), products hi*hi + lo*hi + hi*lo, fp32 accumulate
    (fg_conv_gemm with three K segments);
  * attention, variant "bf16": qkv rounded to bf16, S = q.k fp32, P = exp rounded to bf16, O = P.V / sum, output rounded to
    bf16, projection with att * (w_hi + w_lo)  -- what fg_attention + the x3 projection do today;
  * attention, variant "f32": qkv kept in fp32, softmax(q.k) v in fp32, output split hi + lo, three-segment projection --
    what fg_attention_f32 does.
Reference = the same oracle evaluated in float64.

    python profiles/x3_emulation.py            # tiny golden configs + the full CIFAR-10 network (batch 2)
"""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import edm_oracle as EO  # noqa: E402


def split(x):
    hi = x.float().bfloat16().float()
    lo = (x.float() - hi).bfloat16().float()
    return hi, lo


class Emu:
    def __init__(self, attn):
        self.attn = attn

    def conv2d(self, x, w, padding=0):
        xh, xl = split(x)
        wh, wl = split(w)
        return (self._c(xh, wh, padding) + self._c(xl, wh, padding) + self._c(xh, wl, padding)).float()

    @staticmethod
    def _c(a, b, padding):
        return F_conv2d(a.double(), b.double(), padding=padding)  # exact products, (better than) fp32 accumulation

    def linear(self, sd, k, x):
        xh, xl = split(x)
        wh, wl = split(sd[k + ".weight"])
        y = (xh.double() @ wh.double().t() + xl.double() @ wh.double().t() + xh.double() @ wl.double().t()).float()
        b = sd.get(k + ".bias")
        return y + b.float() if b is not None else y


F_conv2d = F.conv2d


def run(cfg, seed, x_t, t, labels, attn):
    """forward with the emulated rounding points; attn in {"bf16", "f32", None (oracle untouched, float32)}"""
    sd = EO.make_state_dict(cfg, seed)
    with emulated(attn):
        return EO.edm_precond_forward(sd, cfg, x_t, t, labels)


import contextlib  # noqa: E402


@contextlib.contextmanager
def emulated(attn):
    if attn is None:
        yield
        return
    emu = Emu(attn)
    orig_linear, orig_block = EO._linear, EO.unet_block

    def conv_patch(x, w, bias=None, stride=1, padding=0, *a, **kw):
        assert bias is None and stride == 1
        return emu.conv2d(x, w, padding)

    def block(sd_, k, e, x, emb, song):
        if not e["heads"]:
            return orig_block(sd_, k, e, x, emb, song)
        e2 = dict(e, heads=0)
        x = orig_block(sd_, k, e2, x, emb, song)
        eps = 1e-6 if song else 1e-5
        skip_scale = math.sqrt(0.5) if song else 1.0
        B, C, H, W = x.shape
        nh = e["heads"]
        qkv = EO._conv(sd_, k + ".qkv", EO._gn(sd_, k + ".norm2", x, eps)).reshape(B * nh, C // nh, 3, H * W)
        if attn == "bf16":
            qkv = qkv.bfloat16().float()
        q, kk, v = qkv.unbind(2)
        s = torch.einsum("ncq,nck->nqk", q.double(), kk.double()) / math.sqrt(kk.shape[1])
        m = s.amax(dim=2, keepdim=True)
        p = (s - m).float().exp()
        den = p.double().sum(dim=2, keepdim=True)
        if attn == "bf16":
            p = p.bfloat16().float()  # the row sum is taken from the fp32 p, the product from the rounded p
        a = (torch.einsum("nqk,nck->ncq", p.double(), v.double()) / den.transpose(1, 2)).float()
        a = a.reshape(B, C, H, W)
        w = sd_[k + ".proj.weight"]
        if attn == "bf16":
            ah = a.bfloat16().float()
            wh, wl = split(w)
            y = (F_conv2d(ah.double(), wh.double()) + F_conv2d(ah.double(), wl.double())).float()
        else:
            y = emu.conv2d(a, w)
        y = y + sd_[k + ".proj.bias"].reshape(1, -1, 1, 1)
        return (y + x) * skip_scale

    F.conv2d = conv_patch
    EO._linear = emu.linear
    EO.unet_block = block
    try:
        yield
    finally:
        F.conv2d = F_conv2d
        EO._linear = orig_linear
        EO.unet_block = orig_block


def ref64(cfg, seed, x_t, t, labels):
    sd = {k: v.double() for k, v in EO.make_state_dict(cfg, seed).items()}
    return EO.edm_precond_forward(sd, cfg, x_t.double(), t, labels.double() if labels is not None else None)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def golden_checks():
    """every quantity tests/test_edm_parity_gpu.py asserts on the bf16x3 path, predicted by the emulation"""
    from oracle import sampler_oracle as SO

    gold = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    sched = SO.Schedule("edm")
    print(f"{'quantity':34s} {'x3 + bf16 attn':>15s} {'x3 + f32 attn':>14s}   (rel-L2 vs the reference's golden fixture)")
    for name in ("song_tiny", "dhariwal_tiny"):
        g = np.load(os.path.join(gold, f"edm_{name}.npz"))
        cfg, seed = json.loads(str(g["cfg"])), int(g["seed"])
        sd = EO.make_state_dict(cfg, seed)
        T = lambda k: torch.from_numpy(np.asarray(g[k]))  # noqa: E731
        x_t, t, labels, noise = T("x_t"), T("t"), T("labels"), T("noise")
        rows = {}
        for attn in ("bf16", "f32"):
            with emulated(attn):
                x0 = EO.edm_precond_forward(sd, cfg, x_t, t, labels)
                rows.setdefault("fwd_x0", []).append(rel(x0, T("fwd_x0")))
                rows.setdefault("fwd_x0_nolabel", []).append(rel(EO.edm_precond_forward(sd, cfg, x_t, t, None), T("fwd_x0_nolabel")))
                rows.setdefault("fwd_eps", []).append(rel(sched.x0_to_eps(x_t, x0, t), T("fwd_eps")))
                for tag, steps, stype, tl in [("gen1_sde", 1, "sde", None), ("gen2_sde", 2, "sde", None),
                                             ("gen2_ode", 2, "ode", None), ("gen2_sde_tl", 2, "sde", [80.0, 1.1, 0.0]),
                                             ("gen4_sde", 4, "sde", None)]:
                    eps = [T(f"{tag}_eps{i}") for i in range(steps) if f"{tag}_eps{i}" in g]
                    out = SO.generator_fn(lambda x, tb: EO.edm_precond_forward(sd, cfg, x, tb, labels), sched, noise, steps,
                                          tl, eps, stype)
                    rows.setdefault(tag, []).append(rel(out, T(f"{tag}_out")))
        for k, v in rows.items():
            print(f"{name + ' ' + k:34s} {v[0]:15.3e} {v[1]:14.3e}")


def main():
    if "--golden" in sys.argv:
        return golden_checks()
    gold = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    cases = []
    for name in ("song_tiny", "dhariwal_tiny"):
        g = np.load(os.path.join(gold, f"edm_{name}.npz"))
        cfg = json.loads(str(g["cfg"]))
        cases.append((name, cfg, int(g["seed"]), torch.from_numpy(g["x_t"]), torch.from_numpy(g["t"]),
                      torch.from_numpy(g["labels"])))
    if "--full" in sys.argv:
        from bench import CIFAR  # the BASELINE configs[0] network
        torch.manual_seed(1)
        lab = torch.eye(10)[torch.randint(10, (2,))]
        cases.append(("cifar10 full", CIFAR, 0, torch.randn(2, 3, 32, 32) * 80.0, torch.full((2,), 80.0, dtype=torch.float64),
                      lab))
    if "--in64" in sys.argv:
        from bench import IN64  # the BASELINE configs[1] network (295.9 M parameters); one sample, t = 1.1 (second step)
        torch.manual_seed(2)
        lab = torch.eye(1000)[torch.randint(1000, (1,))]
        cases.append(("imagenet64 full", IN64, 0, torch.randn(1, 3, 64, 64) * 1.2, torch.full((1,), 1.1, dtype=torch.float64),
                      lab))
    print(f"{'network':16s} {'fp32 oracle':>12s} {'x3 + bf16 attn':>15s} {'x3 + f32 attn':>14s}   (rel-L2 of x0 vs the float64 oracle)")
    for name, cfg, seed, x_t, t, labels in cases:
        r = ref64(cfg, seed, x_t, t, labels)
        e = [rel(run(cfg, seed, x_t, t, labels, a), r) for a in (None, "bf16", "f32")]
        print(f"{name:16s} {e[0]:12.3e} {e[1]:15.3e} {e[2]:14.3e}")


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    main()
