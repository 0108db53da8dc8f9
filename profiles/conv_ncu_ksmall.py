"""One shape for an ncu source-level capture: the K-small 1x1 qkv projection of the IN-64 UNet (32x32, K=384, N=1152, bf16 out)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import ops  # noqa: E402

n, h, w, cin, cout = 256, 32, 32, 384, 1152
a = torch.randn(n, h, w, cin, device="cuda").bfloat16()
wt = (torch.randn(cout, 1, cin, device="cuda") / math.sqrt(cin)).bfloat16()
b = torch.randn(cout, device="cuda")
out = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.conv_gemm([ops.Seg(a, cin, wt, cin, 1)], n, h, w, cout, out, cout, bias=b)
torch.cuda.synchronize()
