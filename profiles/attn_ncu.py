"""One flash-attention launch per hot shape for `ncu --set full --import-source on -k regex:attention2` (GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import ops  # noqa: E402

dev = "cuda"


def once(b, tq, heads, d, l0, l1):
    dm = heads * d
    q = torch.randn(b, tq, 3 * dm, device=dev).bfloat16()
    segs = []
    if l0:
        kc = torch.randn(b, l0, dm, device=dev).bfloat16()
        segs.append((kc, torch.randn_like(kc), dm, l0 * dm, l0))
    if l1:
        segs.append((q[..., dm:], q[..., 2 * dm:], 3 * dm, tq * 3 * dm, l1))
    out = torch.empty(b, tq, dm, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.flash_attn(q, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    torch.cuda.synchronize()


once(2, 4680, 12, 128, 28080, 4680)   # Wan chunk 6, batch 2: 228 CTAs x 2 (no split)
once(64, 1024, 6, 64, 0, 1024)        # EDM 32x32 / SDXL-like d=64
