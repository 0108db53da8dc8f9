"""Micro-benchmark of the HBM-bound kernels at the sizes of the BASELINE.json configurations (CUDA events, 20 launches
after 3 warm-up launches, working sets larger than the 126 MB L2 unless the configuration itself is smaller -- noted in
the output).  Achieved = algorithmic bytes (every operand read once, every result written once) / time; the peak is the
measured copy bandwidth of MEASURED_PEAKS.json.  usage (GPU box): python profiles/membound_micro.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastgen_b200 import ops  # noqa: E402

dev = "cuda"
PEAK = 6537.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(tag, nbytes, fn):
    try:
        ms = timeit(fn)
        gbs = nbytes / ms / 1e6
        print(f"{tag:74s} {nbytes / 1e6:9.1f} MB {ms * 1000:9.1f} us {gbs:8.0f} GB/s  {gbs / PEAK:5.2f} of measured peak", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{tag:74s} FAILED: {e}", flush=True)


def gn(tag, n, h, w, c, silu, dtype=torch.float32):
    x = torch.randn(n, h, w, c, device=dev).to(dtype)
    gamma, beta = torch.randn(c, device=dev), torch.randn(c, device=dev)
    stats = torch.zeros(n, 32, 2, device=dev, dtype=torch.float64)
    y = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
    ops.gn_stats(x, c, n, h * w, c, 32, stats)
    el = n * h * w * c
    report(f"gn_stats  {tag}", el * x.element_size(), lambda: ops.gn_stats(x, c, n, h * w, c, 32, stats))
    report(f"gn_apply{'+silu' if silu else '     '} {tag}", el * (x.element_size() + 2),
           lambda: ops.gn_apply(x, c, n, h, w, c, 32, stats, 1e-5, gamma, beta, y, c, silu=silu))


def ln(tag, rows, d, dtype, modulate):
    x = (torch.randn(rows, d, device=dev) * 2).to(dtype)
    y = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
    if modulate:
        mod = torch.randn((rows + 1559) // 1560, 2 * d, device=dev)
        kw = dict(scale=mod[:, :d], shift=mod[:, d:], mod_ld=2 * d, rows_per_group=1560)
    else:
        kw = dict(gamma=torch.randn(d, device=dev), beta=torch.randn(d, device=dev))
    report(f"ln_modulate {tag}", rows * d * (x.element_size() + 2), lambda: ops.ln_modulate(x, d, rows, d, 1e-6, y, d, **kw))


def step(tag, n, c, h, w, dtype):
    x_t = torch.randn(n, c, h, w, device=dev).to(dtype)
    eps = torch.randn_like(x_t)
    f = torch.randn(n, h, w, 4, device=dev)
    x0, xn = torch.empty_like(x_t), torch.empty_like(x_t)
    t = torch.full((n,), 80.0, device=dev, dtype=torch.float64)
    tn = torch.full((n,), 1.1, device=dev, dtype=torch.float64)
    cs, co, ones = 0.25 / (t * t + 0.25), t * 0.5 / (t * t + 0.25).sqrt(), torch.ones_like(t)
    es = x_t.element_size()
    el = n * c * h * w
    report(f"edm_step (x0 + re-noise, sde) {tag}", el * (4 + 4 * es),
           lambda: ops.edm_step(f, 4, x_t, x0, cs, co, x_next=xn, eps=eps, alpha_next=ones, sigma_next=tn))


def axpby(tag, shape, dtype):
    a = torch.randn(*shape, device=dev).to(dtype)
    b = torch.randn_like(a)
    out = torch.empty_like(a)
    ca = torch.full((shape[0],), 0.3, device=dev, dtype=torch.float64)
    cb = torch.full((shape[0],), 0.7, device=dev, dtype=torch.float64)
    report(f"axpby (forward_process) {tag}", a.numel() * 3 * a.element_size(), lambda: ops.axpby(a, ca, b, cb, out=out))


print(f"measured copy bandwidth: {PEAK:.0f} GB/s")
# EDM ImageNet-64, batch 256 (configs[1])
gn("IN-64 b256 64x64 c=192 fp32", 256, 64, 64, 192, True)
gn("IN-64 b256 32x32 c=384 fp32", 256, 32, 32, 384, True)
gn("IN-64 b256 64x64 c=384 fp32 (no SiLU)", 256, 64, 64, 384, False)
step("IN-64 b256 [256,3,64,64] fp32 (25 MB: L2-resident)", 256, 3, 64, 64, torch.float32)
step("[4096,3,64,64] fp32 (asymptote)", 4096, 3, 64, 64, torch.float32)
# SDXL batch 16 (configs[2])
gn("SDXL b16 128x128 c=320 fp32", 16, 128, 128, 320, True)
gn("SDXL b16 64x64 c=1920 fp32", 16, 64, 64, 1920, True)
ln("SDXL b16 65536 x 640 fp32 rows, affine", 65536, 640, torch.float32, False)
ln("SDXL b16 16384 x 1280 fp32 rows, affine", 16384, 1280, torch.float32, False)
step("SDXL b16 [16,4,128,128] bf16 (4 MB: L2-resident)", 16, 4, 128, 128, torch.bfloat16)
axpby("SDXL b16 [16,4,128,128] bf16", (16, 4, 128, 128), torch.bfloat16)
# Wan (configs[3], [4])
ln("Wan 1.3B chunk 4680 x 1536 bf16, AdaLN (14 MB: L2-resident)", 4680, 1536, torch.bfloat16, True)
ln("Wan 1.3B full 32760 x 1536 bf16, AdaLN", 32760, 1536, torch.bfloat16, True)
ln("Wan 14B 720p 75600 x 5120 bf16, AdaLN", 75600, 5120, torch.bfloat16, True)
axpby("Wan 480p [1,16,21,60,104] bf16 (4 MB)", (1, 16, 21, 60, 104), torch.bfloat16)
axpby("Wan 720p [1,16,21,90,160] bf16 (10 MB)", (1, 16, 21, 90, 160), torch.bfloat16)
axpby("[64,16,21,60,104] bf16 (asymptote)", (64, 16, 21, 60, 104), torch.bfloat16)
# GEGLU / cast kernels
h = torch.randn(16384, 10240, device=dev).bfloat16()
y = torch.empty(16384, 5120, device=dev, dtype=torch.bfloat16)
report("geglu (stand-alone) 16384 x 2*5120 bf16", 16384 * 5120 * 6, lambda: ops.geglu(h, 16384, 5120, y))
x = torch.randn(16, 64, 64, 640, device=dev)
yb = torch.empty(16, 128, 128, 640, device=dev, dtype=torch.bfloat16)
report("cast_resample up2 SDXL b16 64x64x640 fp32 -> 128x128 bf16", 16 * 64 * 64 * 640 * (4 + 8),
       lambda: ops.cast_resample(x, 640, 16, 64, 64, 640, yb, 640, up2=True))
# the sampler-step rows above run the vector kernels (the default since round 2); `--scalar` adds the scalar forms the
# launcher keeps for unaligned / odd shapes, `--graph` the BASELINE-size calls replayed inside a captured CUDA graph
import sys  # noqa: E402

axpby("[64,16,21,60,104] fp32 (asymptote)", (64, 16, 21, 60, 104), torch.float32)
if "--scalar" in sys.argv:
    ops.sampler_vec(False)
    step("[scalar] [4096,3,64,64] fp32 (asymptote)", 4096, 3, 64, 64, torch.float32)
    axpby("[scalar] [64,16,21,60,104] bf16 (asymptote)", (64, 16, 21, 60, 104), torch.bfloat16)
    ops.sampler_vec(True)
if "--graph" in sys.argv:
    # the BASELINE latents are 4-63 MB: a single eager call is launch latency.  Inside a graph (how the Wan / CIFAR launch
    # lists run them) 20 back-to-back calls on rotating buffers show the kernel time itself.
    def graphed(tag, shape, dtype, bufs=8):
        xs = [torch.randn(shape, device=dev).to(dtype) for _ in range(bufs)]
        es = [torch.randn(shape, device=dev).to(dtype) for _ in range(bufs)]
        a = torch.rand(shape[0], device=dev, dtype=torch.float64)
        b = torch.rand(shape[0], device=dev, dtype=torch.float64)
        outs = [torch.empty_like(xs[0]) for _ in range(bufs)]
        def call(i):
            ops.axpby(xs[i % bufs], a, es[i % bufs], b, out=outs[i % bufs])

        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            call(0)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s_):
                for i in range(20):
                    call(i)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            for _ in range(10):
                g.replay()
            e1.record(s_)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        nbytes = xs[0].numel() * xs[0].element_size() * 3
        print(f"axpby (forward_process) in a CUDA graph, {tag:40s} {nbytes / 1e6:8.1f} MB {us:8.1f} us {nbytes / us / 1e3:8.0f} GB/s  "
              f"{nbytes / us / 1e3 / PEAK:.2f} of measured peak", flush=True)

    graphed("Wan 480p [1,16,21,60,104] bf16", (1, 16, 21, 60, 104), torch.bfloat16)
    graphed("Wan 720p [1,16,21,90,160] bf16", (1, 16, 21, 90, 160), torch.bfloat16)
    graphed("SDXL b16 [16,4,128,128] bf16", (16, 4, 128, 128), torch.bfloat16)
    graphed("IN-64 b256 [256,3,64,64] fp32", (256, 3, 64, 64), torch.float32)


def handoff_lines():
    """SURVEY.md 8f-2 hand-off kernels at the Wan 480p sizes: latent (de)normalisation of a [1,16,21,60,104] latent and the
    clamp of the decoded [1,3,81,480,832] fp32 video (algorithmic bytes: read + write once)"""
    import torch
    from fastgen_b200 import ops

    def timed(f, nbytes, tag, iters=20):
        f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{tag:54s} {ms * 1e3:9.1f} us {nbytes / ms / 1e6:8.1f} GB/s", flush=True)

    a = torch.rand(16, device="cuda") + 0.5
    b = torch.randn(16, device="cuda")
    for bsz in (1, 8):
        lat = torch.randn(bsz, 16, 21, 60, 104, device="cuda").bfloat16()
        timed(lambda: ops.latent_affine(lat, a, b, 0, torch.float32), lat.numel() * 6, f"latent_affine decode bf16->f32 B={bsz} (Wan 480p)")
    vid = torch.randn(1, 3, 81, 480, 832, device="cuda") * 1.2
    timed(lambda: ops.clamp_(vid, -1.0, 1.0), vid.numel() * 8, "clamp_ fp32 [1,3,81,480,832] (decoded video)")
    h = torch.randn(8, 512, 4096, device="cuda")
    lens = torch.randint(5, 300, (8,), device="cuda")
    timed(lambda: ops.zero_tail_rows(h, lens, torch.bfloat16), h.numel() * 6, "zero_tail_rows f32->bf16 [8,512,4096] (prompt embeds)")


if __name__ == "__main__" and "--handoff" in sys.argv:
    handoff_lines()
