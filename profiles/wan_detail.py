"""Wan 2.1 1.3B T2V 480p, 4-step self-forcing (BASELINE configs[3]): per-kernel breakdown of one sample on one GPU.
usage (GPU box): python profiles/wan_detail.py [frames] > gpurun_out/wan_detail.txt"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200.sampler import CausVidModel  # noqa: E402
from fastgen_b200.wan import CausalWan  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 21
B = int(os.environ.get("B", "1"))
torch.manual_seed(0)
t0 = time.time()
with torch.device("cuda"):
    net = CausalWan(total_num_frames=frames, chunk_size=3, use_wan_official_sinusoidal=True, enable_logvar_linear=False)
    torch.nn.init.normal_(net.transformer.proj_out.weight, std=0.02)
net = net.to(torch.bfloat16).eval()
print(f"built in {time.time() - t0:.1f}s, params {sum(p.numel() for p in net.parameters()) / 1e9:.2f} B", flush=True)
noise = torch.randn(B, 16, frames, 60, 104, device="cuda", dtype=torch.bfloat16)
text = torch.randn(B, 512, 4096, device="cuda", dtype=torch.bfloat16)
TL = [0.999, 0.937, 0.833, 0.624, 0.0]


def gen():
    return CausVidModel.generator_fn(net, noise, student_sample_steps=4, t_list=TL, condition=text,
                                     student_sample_type="sde")


gen()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = gen()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"one {frames}-frame sample (batch {B}): {ms:.1f} ms -> {B * 1000.0 / ms:.3f} samples/s; finite={bool(torch.isfinite(out).all())}")
net.profile = {}
gen()
torch.cuda.synchronize()
prof, net.profile = net.profile, None
tot = sum(d["ms"] for d in prof.values())
print(f"profiled total {tot:.1f} ms, {sum(d['launches'] for d in prof.values())} launches")
for name, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    tf = d["flops"] / d["ms"] / 1e9 if d["flops"] else 0.0
    print(f"{d['ms']:9.2f} ms {100 * d['ms'] / tot:5.1f}% x{d['launches']:5d} {name:34s} {tf:8.1f} TFLOP/s")
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
