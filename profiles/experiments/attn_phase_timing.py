"""Clocks per softmax phase of attention2_kernel (experiment build: nvcc ... -DFG_ATTN_TIMING -> libfg_timing.so, see
profiles/attn_r02.md).  usage (GPU box): python profiles/experiments/attn_phase_timing.py [path of the timing library]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fastgen_b200 import _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "experiments", "libfg_timing.so")
from fastgen_b200 import ops  # noqa: E402

dev = "cuda"
NAMES = ["wait S", "TMEM read", "max(+rescale)", "exp + P stores", "st wait + publish"]


def once(tag, b, tq, heads, d, l0, l1):
    dm = heads * d
    q = torch.randn(b, tq, 3 * dm, device=dev).bfloat16()
    segs = []
    if l0:
        kc = torch.randn(b, l0, dm, device=dev).bfloat16()
        segs.append((kc, torch.randn_like(kc), dm, l0 * dm, l0))
    if l1:
        segs.append((q[..., dm:], q[..., 2 * dm:], 3 * dm, tq * 3 * dm, l1))
    out = torch.empty(b, tq, dm, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.flash_attn(q, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.flash_attn(q, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf = (C.c_ulonglong * 16)()
    rc = _lib.lib().fg_debug_attn_timing(buf)
    assert rc == 0, rc
    print(f"{tag}: {ms:.3f} ms, {4.0 * b * tq * (l0 + l1) * dm / ms / 1e9:.0f} TFLOP/s")
    for t in range(2):
        v = [buf[t * 8 + i] for i in range(6)]
        n = max(v[5], 1)
        tot = sum(v[:5])
        print(f"  query tile {t}: {n} key tiles, {tot / n:.0f} clocks / key tile: " +
              ", ".join(f"{NAMES[i]} {v[i] / n:.0f}" for i in range(5)))
    tr = (C.c_longlong * 32)()
    assert _lib.lib().fg_debug_attn_trace(tr) == 0
    base = min(x for x in tr if x > 0)
    for t in range(2):
        for j in range(4):
            a = [tr[(t * 4 + j) * 4 + i] - base for i in range(4)]
            print(f"    tile {t} key tile {j + 2}: P published @{a[0]}, issuer saw it +{a[1] - a[0]}, "
                  f"issued +{a[2] - a[1]}, next S seen +{a[3] - a[2]}  (round trip {a[3] - a[0]})")


once("Wan chunk 6 batch 2, d=128", 2, 4680, 12, 128, 28080, 4680)
once("EDM 32x32 B64 H6 d=64", 64, 1024, 6, 64, 0, 1024)
once("SDXL-like 4096 tokens B4 H10 d=64", 4, 4096, 10, 64, 0, 4096)
