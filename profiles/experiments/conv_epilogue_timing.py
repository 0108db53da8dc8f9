"""Clocks per epilogue phase of conv_gemm_kernel (experiment build: every csrc/*.cu with -DFG_CONV_TIMING -> libfg_timing.so,
see profiles/conv_epilogue_r02.md).  usage (GPU box): python profiles/experiments/conv_epilogue_timing.py [timing library]"""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fastgen_b200 import _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "experiments", "libfg_timing.so")
from fastgen_b200 import ops  # noqa: E402

dev = "cuda"
NAMES = ["wait accumulator", "store-thread prologue", "residual wait", "TMEM read", "tail values", "smem writes + fence",
         "barrier", "store issue", "statistics"]


def once(tag, n, h, w, cin, cout, taps, out_dtype, res=False, stats=False, rowvec=False):
    a = torch.randn(n, h, w, cin, device=dev).bfloat16()
    wt = (torch.randn(cout, taps, cin, device=dev) / math.sqrt(cin * taps)).bfloat16()
    out = torch.empty(n, h, w, cout, device=dev, dtype=out_dtype)
    kw = dict(bias=torch.randn(cout, device=dev))
    if res:
        kw.update(res=torch.randn(n, h, w, cout, device=dev, dtype=out_dtype), res_ld=cout, res_mode=1)
    if rowvec:
        kw.update(rowvec=torch.randn(n, cout, device=dev), rowvec_ld=cout)
    if stats:
        kw.update(stats=torch.zeros(n, 32, 2, device=dev, dtype=torch.float64), stats_groups=32, stats_cg=cout // 32)
    f = lambda: ops.conv_gemm([ops.Seg(a, cin, wt, cin, taps)], n, h, w, cout, out, cout, **kw)  # noqa: E731
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf = (C.c_ulonglong * 24)()
    assert _lib.lib().fg_debug_conv_timing(buf) == 0
    fl = 2.0 * n * h * w * cout * cin * taps
    print(f"{tag}: {ms:.3f} ms, {fl / ms / 1e9:.0f} TFLOP/s")
    for who, name in ((0, "store thread (warp 2)"), (1, "warp 7, lane 0")):
        v = [buf[who * 12 + i] for i in range(12)]
        slabs = max(v[9], 1)
        tot = sum(v[:9]) + v[11]
        print(f"  {name}: {slabs} slabs, {tot / slabs:.0f} clocks / slab: " +
              ", ".join(f"{NAMES[i]} {v[i] / slabs:.0f}" for i in range(9)) + f", tile turnaround {v[11] / slabs:.0f}")


B = 256
if os.environ.get("BISECT"):   # FG_CONV_DBG=16|32|64|128 runs of the experiment build: two shapes only
    print("FG_CONV_DBG =", os.environ.get("FG_CONV_DBG", "0"))
    once("1x1 32x32 K=384 N=384 fp32 + residual + stats", B, 32, 32, 384, 384, 1, torch.float32, res=True, stats=True)
    once("3x3 64x64 K=1728 N=192 fp32 + rowvec + stats", B, 64, 64, 192, 192, 9, torch.float32, stats=True, rowvec=True)
    sys.exit(0)
once("1x1 32x32 K=384 N=384 fp32 + residual + stats", B, 32, 32, 384, 384, 1, torch.float32, res=True, stats=True)
once("1x1 32x32 K=384 N=384 fp32 + residual", B, 32, 32, 384, 384, 1, torch.float32, res=True)
once("1x1 32x32 K=384 N=384 fp32", B, 32, 32, 384, 384, 1, torch.float32)
once("1x1 32x32 K=384 N=1152 bf16 (qkv)", B, 32, 32, 384, 1152, 1, torch.bfloat16)
once("3x3 64x64 K=1728 N=192 fp32 + stats", B, 64, 64, 192, 192, 9, torch.float32, stats=True)
once("3x3 64x64 K=1728 N=192 fp32 + rowvec + stats", B, 64, 64, 192, 192, 9, torch.float32, stats=True, rowvec=True)
once("3x3 64x64 K=1728 N=192 bf16 + stats", B, 64, 64, 192, 192, 9, torch.bfloat16, stats=True)
once("3x3 64x64 K=1728 N=192 fp32 + residual + stats", B, 64, 64, 192, 192, 9, torch.float32, res=True, stats=True)
once("3x3 64x64 K=3456 N=192 fp32 + stats", B, 64, 64, 384, 192, 9, torch.float32, stats=True)
