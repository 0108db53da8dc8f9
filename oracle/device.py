"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): helpers that let the pinned CPU oracles be evaluated ON THE
GPU in true float32 as full-size checkers (the CPU oracles take minutes to hours at the BASELINE
shapes; the same PyTorch code on a B200 in fp32 with TF32 off takes seconds).

  * ``exact_fp32()``   context manager: TF32 off for matmul and cuDNN, float32 matmul precision "highest".
  * ``sdpa(q, k, v, mask)``  what the oracles call for softmax(q k^T / sqrt(d)) v: on the CPU it is
    ``F.scaled_dot_product_attention`` (unchanged: the golden fixtures were generated with it); on CUDA it is an explicit
    matmul / softmax / matmul in the input dtype with fp32 accumulation, chunked over query rows so that a 32 760-token
    sequence needs a few GB, and free of the TF32 tensor-core path PyTorch's fused fp32 attention kernels may take.
Nothing under fastgen_b200/ imports this module."""
from __future__ import annotations

import math
from contextlib import contextmanager
from typing import Optional

import torch
import torch.nn.functional as F


@contextmanager
def exact_fp32():
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.get_float32_matmul_precision())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev[0], prev[1]
        torch.set_float32_matmul_precision(prev[2])


def exact_sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_mask: Optional[torch.Tensor] = None,
               max_score_bytes: int = 4 << 30) -> torch.Tensor:
    """q [B, H, Tq, d], k / v [B, H, Tk, d]; attn_mask: None, boolean [Tq, Tk] (True = attend), or a floating additive bias
    broadcastable to [B, H, Tq, Tk] with a singleton query dim (the SD 1.5 `encoder_attention_mask` bias [B, 1, 1, Tk]).
    Scores, softmax and the P V product in float32 whatever the input dtype; result in q.dtype."""
    B, H, Tq, d = q.shape
    Tk = k.shape[2]
    scale = 1.0 / math.sqrt(d)
    rows = max(1, min(Tq, max_score_bytes // max(1, B * H * Tk * 4)))
    out = torch.empty_like(q)
    kf, vf = k.float(), v.float()
    for r0 in range(0, Tq, rows):
        r1 = min(Tq, r0 + rows)
        s = torch.matmul(q[:, :, r0:r1].float(), kf.transpose(-1, -2)) * scale
        if attn_mask is not None and attn_mask.is_floating_point():
            s = s + attn_mask.to(device=s.device, dtype=s.dtype)
        elif attn_mask is not None:
            s = s.masked_fill(~attn_mask[r0:r1].to(s.device), float("-inf"))
        out[:, :, r0:r1] = torch.matmul(torch.softmax(s, dim=-1), vf).to(q.dtype)
        del s
    return out


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    if q.is_cuda:
        return exact_sdpa(q, k, v, attn_mask)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask)
