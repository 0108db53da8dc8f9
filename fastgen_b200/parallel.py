"""Host-side partitioning logic for the multi-GPU paths (one process per GPU, ``torch.distributed``).

* image models (EDM / SDXL / DiT): independent samples -> the batch list is dealt round-robin to the ranks exactly like
  the reference's FID script (``rank_batches = all_batches[rank::world]``, scripts/compute_fid_from_ckpts.py:88-93),
  per-rank seed ``seed + rank`` (utils/basic_utils.py:140-143); no collective on the data path.
* Wan video: Ulysses sequence parallelism.  A rank owns tokens ``[rank*fs/P, (rank+1)*fs/P)`` of EVERY frame (fs =
  tokens per frame), all heads; around each self-attention an all-to-all turns that into (all tokens, H/P heads) and
  back.  The descriptors below are the 4-D strided copies (``fg_strided_copy``) that pack / unpack the exchange
  buffers; they are plain integer arithmetic so the layout can be unit-tested on CPU with the gloo backend.

A descriptor is ``dict(n=[4 extents], ss=[4 src strides], ds=[4 dst strides], item=elements per item, src_off=, dst_off=)``
with strides / offsets in elements.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import torch


def rank_batches(all_batches: Sequence[Any], rank: int, world: int) -> List[Any]:
    return list(all_batches[rank::world])


def rank_seed(seed: int, rank: int) -> int:
    return seed + rank


def shard_rows_desc(B: int, F: int, fs: int, P: int, rank: int, D: int) -> Dict[str, Any]:
    """global token rows [(b, f, t)] -> this rank's rows [(b, f, i)], t = rank*fs/P + i."""
    fs_l = fs // P
    return dict(n=[B, F, fs_l, 1], ss=[F * fs * D, fs * D, D, 0], ds=[F * fs_l * D, fs_l * D, D, 0], item=D,
                src_off=rank * fs_l * D, dst_off=0)


def a2a_pack_desc(n_l: int, P: int, D: int, Dl: int) -> Dict[str, Any]:
    """one sample: qkv[r][j][dst][Dl] (row = [q | k | v] x all heads) -> send[dst][r][j][Dl]."""
    return dict(n=[n_l, 3, P, 1], ss=[3 * D, D, Dl, 0], ds=[3 * Dl, Dl, n_l * 3 * Dl, 0], item=Dl, src_off=0, dst_off=0)


def a2a_unpack_desc(B: int, P: int, n_l: int, D: int, Dl: int) -> Dict[str, Any]:
    """recv[b][src][r][Dl] (attention output of src's heads for my rows) -> att[b][r][src][Dl] = rows x all heads."""
    return dict(n=[B, P, n_l, 1], ss=[P * n_l * Dl, n_l * Dl, Dl, 0], ds=[n_l * D, Dl, D, 0], item=Dl, src_off=0, dst_off=0)


def unshard_desc(P: int, B: int, F: int, fs: int, kout: int) -> Dict[str, Any]:
    """all-gathered per-rank outputs [src][(b, f)][i] -> natural token order [(b, f)][src*fs/P + i]."""
    fs_l = fs // P
    return dict(n=[P, B * F, fs_l, 1], ss=[B * F * fs_l * kout, fs_l * kout, kout, 0],
                ds=[fs_l * kout, fs * kout, kout, 0], item=kout, src_off=0, dst_off=0)


def peer_put_qkv_desc(B: int, n: int, P: int, rank: int, D: int, Dl: int) -> Dict[str, Any]:
    """pack + exchange in one pass (``fg_peer_put``): this rank's qkv[b][r][j][dst][Dl] -> rank dst's
    full[b][rank*n/P + r][j][Dl].  Same result as a2a_pack_desc + all_to_all_single on every destination."""
    n_l = n // P
    return dict(n=[B, n_l, 3, 1], ss=[n_l * 3 * D, 3 * D, D, 0], ds=[n * 3 * Dl, 3 * Dl, Dl, 0], item=Dl,
                src_peer_stride=Dl, dst_off=rank * n_l * 3 * Dl)


def peer_put_out_desc(B: int, n: int, P: int, rank: int, D: int, Dl: int) -> Dict[str, Any]:
    """exchange + unpack in one pass: this rank's att_full[b][dst*n/P + r][Dl] (its heads, all tokens) -> rank dst's
    att[b][r][rank*Dl : (rank+1)*Dl].  Same result as all_to_all_single + a2a_unpack_desc on every destination."""
    n_l = n // P
    return dict(n=[B, n_l, 1, 1], ss=[n * Dl, Dl, 0, 0], ds=[n_l * D, D, 0, 0], item=Dl, src_peer_stride=n_l * Dl,
                dst_off=rank * Dl)


def peer_gather_desc(lrows: int, kout: int, rank: int) -> Dict[str, Any]:
    """all-gather as a put: this rank's [lrows, kout] rows -> slot ``rank`` of every rank's gathered [P, lrows, kout]
    (the same source for every destination: src_peer_stride = 0)."""
    return dict(n=[lrows, 1, 1, 1], ss=[kout, 0, 0, 0], ds=[kout, 0, 0, 0], item=kout, src_peer_stride=0,
                dst_off=rank * lrows * kout)


def apply_peer_put_cpu(desc: Dict[str, Any], src: torch.Tensor, dsts: Sequence[torch.Tensor]) -> None:
    """Reference execution of a peer-put descriptor: ``dsts[q]`` is rank q's destination buffer."""
    n, item = desc["n"], desc["item"]
    for q, dst in enumerate(dsts):
        assert src.is_contiguous() and dst.is_contiguous()
        sv = torch.as_strided(src, n + [item], desc["ss"] + [1], storage_offset=src.storage_offset() + q * desc["src_peer_stride"])
        dv = torch.as_strided(dst, n + [item], desc["ds"] + [1], storage_offset=dst.storage_offset() + desc["dst_off"])
        dv.copy_(sv)


def apply_desc_cpu(desc: Dict[str, Any], src: torch.Tensor, dst: torch.Tensor) -> None:
    """Reference execution of a descriptor with torch.as_strided (CPU tests of the layout; the product path runs the
    same descriptor through ``fg_strided_copy``)."""
    n, item = desc["n"], desc["item"]
    assert src.is_contiguous() and dst.is_contiguous()
    sv = torch.as_strided(src, n + [item], desc["ss"] + [1], storage_offset=src.storage_offset() + desc["src_off"])
    dv = torch.as_strided(dst, n + [item], desc["ds"] + [1], storage_offset=dst.storage_offset() + desc["dst_off"])
    dv.copy_(sv)
