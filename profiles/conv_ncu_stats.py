"""ncu source-level capture: 1x1 projection + fp32 residual + fused GroupNorm statistics (32x32, K=384, N=384), the
epilogue-heaviest layer type of the IN-64 UNet.  STATS=0 drops the statistics for an A/B timing."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import ops  # noqa: E402

n, h, w, cin, cout = 256, 32, 32, 384, 384
a = torch.randn(n, h, w, cin, device="cuda").bfloat16()
wt = (torch.randn(cout, 1, cin, device="cuda") / math.sqrt(cin)).bfloat16()
b = torch.randn(cout, device="cuda")
res = torch.randn(n, h, w, cout, device="cuda")
out = torch.empty(n, h, w, cout, device="cuda")
stats = torch.zeros(n, 32, 2, device="cuda", dtype=torch.float64)
kw = dict(bias=b, res=res, res_ld=cout, res_mode=1)
if os.environ.get("STATS", "1") == "1":
    kw.update(stats=stats, stats_groups=32, stats_cg=cout // 32, stats_ch0=0)
f = lambda: ops.conv_gemm([ops.Seg(a, cin, wt, cin, 1)], n, h, w, cout, out, cout, **kw)  # noqa: E731
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f()
e1.record()
torch.cuda.synchronize()
print(f"stats={os.environ.get('STATS', '1')}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
