"""One shape for an ncu capture: the ffn2 GEMM of a Wan 1.3B block on a 4-way sequence-parallel shard (1170 rows)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import ops  # noqa: E402

rows, k, cout = int(os.environ.get("ROWS", 1170)), 8960, 1536
a = torch.randn(rows, k, device="cuda").bfloat16()
wt = (torch.randn(cout, k, device="cuda") / math.sqrt(k)).bfloat16()
b = torch.randn(cout, device="cuda")
out = torch.zeros(rows, cout, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.conv_gemm([ops.Seg(a, k, wt, k, 1)], rows, 1, 1, cout, out, cout, bias=b, res=out, res_ld=cout, res_mode=1)
torch.cuda.synchronize()
