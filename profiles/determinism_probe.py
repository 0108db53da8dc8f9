"""Run-to-run and batch-composition variation of the IN-64 network on the GPU (bf16 path): the numbers behind the bounds of
tests/test_edm_parity_gpu.py::test_batch_independence_at_bench_batch.  GroupNorm statistics are merged with fp64 atomics
(order varies run to run at the 1e-16 level); once such a difference flips one bf16 rounding it propagates."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200.edm import EDMPrecond  # noqa: E402
from oracle import edm_oracle as EO  # noqa: E402

IN64 = dict(img_resolution=64, img_channels=3, label_dim=1000, sigma_shift=0.0, sigma_data=0.5,
            model_type="DhariwalUNet", augment_dim=0, model_channels=192, channel_mult=[1, 2, 3, 4], channel_mult_emb=4,
            num_blocks=3, attn_resolutions=[32, 16, 8], dropout=0.0, label_dropout=0, r_timestep=False, drop_precond=None)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


for compute in ("bf16", "bf16x3"):
    net = EDMPrecond(compute=compute, **IN64)
    net.load_state_dict(EO.make_state_dict(IN64, 0), strict=False)
    net = net.cuda().eval()
    torch.manual_seed(2)
    B = 64
    x = torch.randn(B, 3, 64, 64, device="cuda") * 80
    t = torch.full((B,), 80.0, device="cuda", dtype=torch.float64)
    lab = torch.eye(1000, device="cuda")[torch.randint(1000, (B,), device="cuda")]
    base = net(x, t, condition=lab, fwd_pred_type="x0").clone()
    runs = [rel(net(x, t, condition=lab, fwd_pred_type="x0"), base) for _ in range(8)]
    bitwise = sum(r == 0.0 for r in runs)
    small = [rel(base[lo:lo + 2], net(x[lo:lo + 2].clone(), t[:2], condition=lab[lo:lo + 2], fwd_pred_type="x0"))
             for lo in (0, 10, 30, 62)]
    print(f"{compute}: run-to-run rel-L2 max {max(runs):.3e} ({bitwise}/8 bit-identical); "
          f"sample in a batch of 64 vs in a batch of 2: max {max(small):.3e}", flush=True)
    del net
