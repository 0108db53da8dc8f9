"""ncu --set full target: two K-large 3x3 convolutions of the IN-64 UNet at batch 256 with the CTA-pair kernels
(32x32 K=3456 N=384 and 64x64 K=1728 N=192, fp32 out)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import ops  # noqa: E402

for (h, cin, cout) in ((32, 384, 384), (64, 192, 192)):
    n = 256
    a = torch.randn(n, h, h, cin, device="cuda").bfloat16()
    wt = (torch.randn(cout, 9, cin, device="cuda") / math.sqrt(9 * cin)).bfloat16()
    b = torch.randn(cout, device="cuda")
    out = torch.empty(n, h, h, cout, device="cuda")
    for _ in range(2):
        ops.conv_gemm([ops.Seg(a, cin, wt, cin, 9)], n, h, h, cout, out, cout, bias=b)
torch.cuda.synchronize()
