"""Measured parity margins of the EDM networks at full width and depth against the pinned oracle, with the conv0 -> norm1
tensor in bf16 (default) and in fp32 (FASTGEN_B200_BF16_H1=0).  usage (GPU box): python profiles/parity_margins_r02.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_edm_parity_gpu as T  # noqa: E402
from oracle import edm_oracle as EO  # noqa: E402
from oracle import sampler_oracle as SO  # noqa: E402


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


for name, cfg, batch, tvals in (("cifar10", T.CIFAR, 4, [SO.Schedule("edm").get_t_list(1).tolist()[0]]),
                                ("imagenet64", T.IN64, 2, [80.0, 1.1])):
    torch.manual_seed(1)
    R, L = cfg["img_resolution"], cfg["label_dim"]
    x = torch.randn(batch, 3, R, R)
    labels = torch.eye(L)[torch.randint(L, (batch,))]
    for h1 in (True, False):
        net, sd = T._net(cfg, 0, "bf16")
        net.bf16_h1 = h1
        net._invalidate()
        for tv in tvals:
            xt = x * tv
            t = torch.full((batch,), tv, dtype=torch.float64)
            with torch.no_grad():
                ref = EO.edm_precond_forward(sd, cfg, xt, t, labels)
            out = net(xt.cuda(), t.cuda(), condition=labels.cuda(), fwd_pred_type="x0")
            print(f"{name:10s} t={tv:7.3f} conv0->norm1 {'bf16' if h1 else 'fp32'}: x0 rel-L2 {rel(out, ref):.3e}  (bound 1e-2)", flush=True)
