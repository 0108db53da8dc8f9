"""Per-launch breakdown of one ImageNet-64 forward at batch 256 (CUDA events around every launch).
usage (GPU box): python profiles/detail.py > gpurun_out/detail.txt"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import IN64, _randomize_zero_init_  # noqa: E402
from fastgen_b200.edm import EDMPrecond  # noqa: E402

B = int(os.environ.get("B", "256"))
net = EDMPrecond(**IN64)
_randomize_zero_init_(net, 0)
net = net.cuda().eval()
x = torch.randn(B, 3, 64, 64, device="cuda") * 80
t = torch.full((B,), 80.0, device="cuda", dtype=torch.float64)
lab = torch.eye(1000, device="cuda")[torch.randint(1000, (B,), device="cuda")]
for _ in range(2):
    net(x, t, condition=lab)
net.profile, net.profile_detail = {}, []
net(x, t, condition=lab)
agg = collections.OrderedDict()
for name, desc, ms, fl, nb in net.profile_detail:
    d = agg.setdefault((name, desc), [0, 0.0, 0.0, 0.0])
    d[0] += 1
    d[1] += ms
    d[2] += fl
    d[3] += nb
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.2f} ms for one forward, batch {B}")
for (name, desc), (cnt, ms, fl, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}% x{cnt:3d} {name[3:]:12s} {desc:45s} "
          f"{fl / ms / 1e9 if fl else 0:8.1f} TFLOP/s {nb / ms / 1e6:8.1f} GB/s")
