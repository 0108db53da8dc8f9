"""Micro-benchmark of fg_flash_attn on the hot-path shapes (CUDA events).  usage (GPU box): python profiles/attn_micro.py
Environment switches for A/B runs: FG_ATTN_V2=0 (first-generation kernel), FG_ATTN_POLY=0|2|3|4, FG_ATTN_SPLIT=n."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import ops  # noqa: E402

dev = "cuda"


def run(tag, b, tq, heads, d, l0, l1, ws=True, iters=10):
    dm = heads * d
    q = torch.randn(b, tq, 3 * dm, device=dev).bfloat16()
    segs = []
    if l0:
        kc = torch.randn(b, l0, dm, device=dev).bfloat16()
        vc = torch.randn(b, l0, dm, device=dev).bfloat16()
        segs.append((kc, vc, dm, l0 * dm, l0))
    if l1:
        segs.append((q[..., dm:], q[..., 2 * dm:], 3 * dm, tq * 3 * dm, l1))
    out = torch.empty(b, tq, dm, device=dev, dtype=torch.bfloat16)
    p = ops.flash_attn_params(q, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    w = torch.empty(max(ops.flash_attn_workspace_size(p), 16), device=dev, dtype=torch.uint8) if ws else None
    f = lambda: ops.flash_attn(q, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out, dm, tq * dm, workspace=w)  # noqa: E731
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * b * tq * (l0 + l1) * dm
    print(f"{tag:44s} {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


run("wan chunk0 self  tq4680 kv4680", 1, 4680, 12, 128, 0, 4680)
run("wan chunk3 self  tq4680 kv14040+4680", 1, 4680, 12, 128, 14040, 4680)
run("wan chunk6 self  tq4680 kv28080+4680", 1, 4680, 12, 128, 28080, 4680)
run("wan chunk6 self  (no split workspace)", 1, 4680, 12, 128, 28080, 4680, ws=False)
run("wan B=2 chunk6 self", 2, 4680, 12, 128, 28080, 4680)
run("wan cross tq4680 kv512", 1, 4680, 12, 128, 512, 0)
run("wan SP4 chunk6 (3 heads)", 1, 4680, 3, 128, 28080, 4680)
run("wan bidir full tq32760", 1, 32760, 12, 128, 0, 32760, iters=2)
run("edm 32x32 B256 H6 d64", 256, 1024, 6, 64, 0, 1024)
run("edm 16x16 B256 H9 d64", 256, 256, 9, 64, 0, 256)


def run_f32(tag, b, tq, heads, d, iters=5):
    """fg_attention_f32 (the fp32 CUDA-core attention of the compute="bf16x3" parity path): `--f32` adds these lines"""
    dm = heads * d
    qkv = torch.randn(b, tq, 3 * dm, device=dev)
    hi = torch.empty(b, tq, dm, device=dev, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    f = lambda: ops.attention_f32(qkv, 3 * dm, b, tq, heads, d, hi, lo, dm)  # noqa: E731
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{tag:44s} {ms:8.3f} ms {4.0 * b * tq * tq * dm / ms / 1e9:8.1f} TFLOP/s (fp32 FMA)", flush=True)


if "--f32" in sys.argv:
    run_f32("f32 edm 32x32 B256 H6 d64", 256, 1024, 6, 64)
    run_f32("f32 edm 16x16 B256 H9 d64", 256, 256, 9, 64)
    run_f32("f32 edm 8x8 B256 H12 d64", 256, 64, 12, 64)
    run_f32("f32 cifar 16x16 B64 H1 d256", 64, 256, 1, 256)
