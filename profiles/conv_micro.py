"""Micro-benchmark of fg_conv_gemm on single layer shapes (CUDA events, L2 flushed between runs by the sheer
tensor sizes).  usage (GPU box): python profiles/conv_micro.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import ops  # noqa: E402

dev = "cuda"


def run(tag, n, h, w, cin, cout, taps, out_dtype, res=False, bias=True, iters=5):
    a = torch.randn(n, h, w, cin, device=dev).bfloat16()
    wt = (torch.randn(cout, taps, cin, device=dev) / math.sqrt(cin * taps)).bfloat16()
    b = torch.randn(cout, device=dev) if bias else None
    out = torch.empty(n, h, w, cout, device=dev, dtype=out_dtype)
    r = torch.randn(n, h, w, cout, device=dev) if res else None
    kw = dict(bias=b)
    if res:
        kw.update(res=r, res_ld=cout, res_mode=1)
    f = lambda: ops.conv_gemm([ops.Seg(a, cin, wt, cin, taps)], n, h, w, cout, out, cout, **kw)  # noqa: E731
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * h * w * cout * cin * taps
    by = n * h * w * (cin * 2 + cout * out.element_size() + (cout * 4 if res else 0))
    print(f"{tag:34s} {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s {by / ms / 1e6:8.1f} GB/s", flush=True)


def run_linear(tag, rows, k, cout, iters=20):
    """token-row GEMM of the Wan blocks (bf16 rows in, bf16 out + bias + residual): rows = tokens of a rank's shard"""
    a = torch.randn(rows, k, device=dev).bfloat16()
    wt = (torch.randn(cout, k, device=dev) / math.sqrt(k)).bfloat16()
    b = torch.randn(cout, device=dev)
    out = torch.zeros(rows, cout, device=dev, dtype=torch.bfloat16)
    f = lambda: ops.conv_gemm([ops.Seg(a, k, wt, k, 1)], rows, 1, 1, cout, out, cout, bias=b, res=out, res_ld=cout, res_mode=1)  # noqa: E731
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{tag:34s} rows {rows:5d} {ms * 1e3:8.1f} us {2.0 * rows * k * cout / ms / 1e9:8.1f} TFLOP/s", flush=True)


B = 256
if len(sys.argv) > 1 and sys.argv[1] == "--shard":   # Wan 1.3B GEMMs on sequence-parallel shards (FG_CONV_BN=128|192|256)
    for rows in (4680, 2340, 1170):
        run_linear("qkv  K1536 N4608", rows, 1536, 4608)
        run_linear("o1   K1536 N1536", rows, 1536, 1536)
        run_linear("ffn1 K1536 N8960", rows, 1536, 8960)
        run_linear("ffn2 K8960 N1536", rows, 8960, 1536)
    sys.exit(0)
if len(sys.argv) > 1:  # short list for A/B experiments (FG_CONV_* environment switches)
    run("qkv 32x32 K384 N1152 bf16", B, 32, 32, 384, 1152, 1, torch.bfloat16)
    run("proj 32x32 K384 N384 f32", B, 32, 32, 384, 384, 1, torch.float32)
    run("3x3 64x64 K1728 N192 f32", B, 64, 64, 192, 192, 9, torch.float32)
    run("3x3 32x32 K3456 N384 f32", B, 32, 32, 384, 384, 9, torch.float32)
    sys.exit(0)
run("qkv 32x32 K384 N1152 bf16", B, 32, 32, 384, 1152, 1, torch.bfloat16)
run("qkv 32x32 K384 N1152 f32", B, 32, 32, 384, 1152, 1, torch.float32)
run("qkv nobias bf16", B, 32, 32, 384, 1152, 1, torch.bfloat16, bias=False)
run("proj 32x32 K384 N384 f32 +res", B, 32, 32, 384, 384, 1, torch.float32, res=True)
run("proj 32x32 K384 N384 f32", B, 32, 32, 384, 384, 1, torch.float32)
run("proj 32x32 K384 N384 bf16", B, 32, 32, 384, 384, 1, torch.bfloat16)
run("3x3 32x32 K3456 N384 f32", B, 32, 32, 384, 384, 9, torch.float32)
run("3x3 32x32 K3456 N384 f32 +res", B, 32, 32, 384, 384, 9, torch.float32, res=True)
run("3x3 64x64 K1728 N192 f32", B, 64, 64, 192, 192, 9, torch.float32)
run("3x3 64x64 K1728 N192 f32 +res", B, 64, 64, 192, 192, 9, torch.float32, res=True)
run("3x3 64x64 K1728 N192 bf16", B, 64, 64, 192, 192, 9, torch.bfloat16)
run("1x1 64x64 K192 N192 bf16", B, 64, 64, 192, 192, 1, torch.bfloat16)
run("1x1 8x8 K768 N2304 bf16", B, 8, 8, 768, 2304, 1, torch.bfloat16)
