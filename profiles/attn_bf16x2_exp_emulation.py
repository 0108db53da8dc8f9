"""CPU pre-study for the d = 64 attention (profiles/attn_r01.md: the softmax instruction stream bounds it): what would
`ex2.approx.ftz.bf16x2` (two exponentials per MUFU issue, the result already packed bf16) cost in accuracy?

Emulates the rounding points of flash attention with bf16 operands: S in fp32, x = S*scale*log2e - rowmax in fp32, then
  current kernel : p = bf16(exp2(x))            (fp32 exponential, P rounded to bf16 for the P.V product; row sum of fp32 p)
  bf16x2 variant : p = bf16(exp2(bf16(x)))      (the ARGUMENT is rounded to bf16 as well; row sum of the same p)
  mixed variant  : bf16x2 for a share of the columns only (like the polynomial share today)
and reports rel-L2 of the output against float64 attention on the same bf16 q, k, v."""
import math

import torch

torch.manual_seed(0)


def attn(q, k, v, mode, share=1.0):
    d = q.shape[-1]
    s = (q.double() @ k.double().transpose(-1, -2)).float()
    x = s * (1.4426950408889634 / math.sqrt(d))
    x = x - x.amax(dim=-1, keepdim=True)
    if mode == "ref":
        p = torch.exp2(x.double())
        return (p @ v.double()) / p.sum(-1, keepdim=True)
    p32 = torch.exp2(x)
    if mode == "current":
        p, den = p32.bfloat16().float(), p32.sum(-1, keepdim=True)
    elif mode == "ones_column":      # row sum taken by the P.V product itself (a ones column in V): sum of the ROUNDED p
        p = p32.bfloat16().float()
        den = p.sum(-1, keepdim=True)
    else:
        pb = torch.exp2(x.bfloat16().float()).bfloat16().float()
        if share < 1.0:
            cols = torch.arange(x.shape[-1]) % 8 < round(8 * share)
            p = torch.where(cols, pb, p32.bfloat16().float())
            den = torch.where(cols, pb, p32).sum(-1, keepdim=True)
        else:
            p, den = pb, pb.sum(-1, keepdim=True)   # the sum of the SAME rounded weights: systematic bias cancels
    o = (p.double() @ v.double()) / den.double()
    return o.float().bfloat16().double()


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


print(f"{'case':44s} {'current':>10s} {'bf16x2 all':>11s} {'bf16x2 1/2':>11s} {'bf16x2 1/4':>11s} {'ones col.':>11s}")
for name, T, d, qs in [("N(0,1) q,k  T=1024 d=64 (EDM 32x32)", 1024, 64, 1.0), ("sharp (q,k x2) T=1024 d=64", 1024, 64, 2.0),
                       ("very sharp (x4) T=1024 d=64", 1024, 64, 4.0), ("N(0,1) T=4096 d=64 (SDXL 64x64)", 4096, 64, 1.0),
                       ("N(0,1) T=4680 d=128 (Wan chunk)", 4680, 128, 1.0)]:
    q = (torch.randn(4, T, d) * qs).bfloat16()
    k = (torch.randn(4, T, d) * qs).bfloat16()
    v = torch.randn(4, T, d).bfloat16()
    r = attn(q, k, v, "ref")
    row = [rel(attn(q, k, v, "current"), r), rel(attn(q, k, v, "bf16x2"), r), rel(attn(q, k, v, "bf16x2", 0.5), r),
           rel(attn(q, k, v, "bf16x2", 0.25), r), rel(attn(q, k, v, "ones_column"), r)]
    print(f"{name:44s} " + " ".join(f"{e:11.3e}" for e in row))
