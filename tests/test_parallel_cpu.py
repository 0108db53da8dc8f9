"""N > 1 host logic on CPU (gloo, world_size 2): batch sharding of the image samplers and the Ulysses exchange layout
of the Wan path (pack -> all_to_all -> attention over (all tokens, local heads) -> all_to_all -> unpack)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from fastgen_b200 import parallel as PL


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_rank_batches_partition_like_the_reference():
    batches = list(range(11))
    parts = [PL.rank_batches(batches, r, 4) for r in range(4)]
    assert parts[0] == [0, 4, 8] and parts[3] == [3, 7]
    assert sorted(sum(parts, [])) == batches
    assert PL.rank_seed(7, 3) == 10


def _ulysses_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical "global" tensors on every rank
        B, Fr, fs, H, d = 2, 3, 8, 4, 16
        D, P = H * d, world
        n, fs_l = Fr * fs, fs // P
        n_l, Hl = n // P, H // P
        Dl = Hl * d
        hs_full = torch.randn(B * n, D)
        qkv_full = torch.randn(B * n, 3 * D)          # what every rank WOULD compute for all tokens
        # 1. token shard of this rank
        hs = torch.zeros(B * n_l, D)
        PL.apply_desc_cpu(PL.shard_rows_desc(B, Fr, fs, P, rank, D), hs_full, hs)
        idx = torch.arange(B * n).view(B, Fr, fs)[:, :, rank * fs_l:(rank + 1) * fs_l].reshape(-1)
        assert torch.equal(hs, hs_full[idx])
        qkv = qkv_full[idx].contiguous()              # local rows, all heads
        # 2. pack + all-to-all: (local tokens, all heads) -> (all tokens, local heads)
        send = torch.zeros(B, P, n_l, 3 * Dl)
        full = torch.zeros(B, n, 3 * Dl)
        for b in range(B):
            PL.apply_desc_cpu(PL.a2a_pack_desc(n_l, P, D, Dl), qkv[b * n_l:(b + 1) * n_l], send[b])
            dist.all_to_all_single(full[b], send[b])
        # gathered token order of sample b: (src, f, i); heads [rank*Hl, (rank+1)*Hl)
        order = torch.arange(n).view(Fr, P, fs_l).permute(1, 0, 2).reshape(-1)
        for b in range(B):
            want = qkv_full[b * n:(b + 1) * n][order].view(n, 3, H, d)[:, :, rank * Hl:(rank + 1) * Hl].reshape(n, 3 * Dl)
            assert torch.equal(full[b], want)
        # 3. attention on local heads (keys in the gathered order: softmax is permutation invariant)
        q, k, v = (full.view(B, n, 3, Hl, d)[:, :, j].transpose(1, 2) for j in range(3))
        attf = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, n, Dl).contiguous()
        # 4. all-to-all back + unpack: (all tokens, local heads) -> (local tokens, all heads)
        recv = torch.zeros(B, P, n_l, Dl)
        for b in range(B):
            dist.all_to_all_single(recv[b], attf[b])
        att = torch.zeros(B * n_l, D)
        PL.apply_desc_cpu(PL.a2a_unpack_desc(B, P, n_l, D, Dl), recv, att)
        qa, ka, va = (qkv_full.view(B, n, 3, H, d)[:, :, j].transpose(1, 2) for j in range(3))
        ref = F.scaled_dot_product_attention(qa, ka, va).transpose(1, 2).reshape(B * n, D)
        assert torch.allclose(att, ref[idx], atol=1e-5)
        # 5. gather the per-rank outputs back into natural token order
        kout = 8
        out_full = torch.randn(B * n, kout)
        parts = [torch.zeros(B * n_l, kout) for _ in range(P)]
        dist.all_gather(parts, out_full[idx].contiguous())   # (gloo has no all_gather_into_tensor)
        gathered = torch.stack(parts, 0)
        nat = torch.zeros(B * n, kout)
        PL.apply_desc_cpu(PL.unshard_desc(P, B, Fr, fs, kout), gathered, nat)
        assert torch.equal(nat, out_full)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_ulysses_layout_world_size_2():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ulysses_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world))


def test_peer_put_descriptors_equal_pack_all_to_all_unpack():
    """the one-pass exchange (fg_peer_put: every rank stores its slices straight into the peers' buffers) must leave
    exactly what pack -> all_to_all_single -> (attention) -> all_to_all_single -> unpack leaves; all P ranks are
    simulated in one process, the all-to-all being the block transpose send[src][dst] -> recv[dst][src]"""
    import pytest

    for P in (2, 4, 8):
        torch.manual_seed(P)
        B, Fr, fs, H, d = 2, 3, 16, 8, 8
        D = H * d
        n, n_l, Dl = Fr * fs, Fr * fs // P, (H // P) * d
        qkv = [torch.randn(B * n_l, 3 * D) for _ in range(P)]            # local rows x all heads, per rank
        # NCCL path: pack, exchange, and for the way back exchange, unpack
        send = [torch.zeros(B, P, n_l, 3 * Dl) for _ in range(P)]
        for r in range(P):
            for b in range(B):
                PL.apply_desc_cpu(PL.a2a_pack_desc(n_l, P, D, Dl), qkv[r][b * n_l:(b + 1) * n_l], send[r][b])
        full_ref = [torch.stack([torch.cat([send[src][b, dst] for src in range(P)], 0) for b in range(B)], 0)
                    for dst in range(P)]                                   # [B, n, 3*Dl] per rank
        # peer path, direction 1
        full = [torch.full((B, n, 3 * Dl), float("nan")) for _ in range(P)]
        for r in range(P):
            PL.apply_peer_put_cpu(PL.peer_put_qkv_desc(B, n, P, r, D, Dl), qkv[r], full)
        for r in range(P):
            assert torch.equal(full[r], full_ref[r]), (P, r)
        # direction 2: some per-rank attention output over (all tokens, local heads)
        attf = [torch.randn(B, n, Dl) for _ in range(P)]
        att_ref = []
        for dst in range(P):
            recv = torch.stack([torch.stack([attf[src][b, dst * n_l:(dst + 1) * n_l] for src in range(P)], 0)
                                for b in range(B)], 0)                     # [B, P, n_l, Dl]
            a = torch.zeros(B * n_l, D)
            PL.apply_desc_cpu(PL.a2a_unpack_desc(B, P, n_l, D, Dl), recv.contiguous(), a)
            att_ref.append(a)
        att = [torch.full((B * n_l, D), float("nan")) for _ in range(P)]
        for r in range(P):
            PL.apply_peer_put_cpu(PL.peer_put_out_desc(B, n, P, r, D, Dl), attf[r], att)
        for r in range(P):
            assert torch.equal(att[r], att_ref[r]), (P, r)
        # the final all-gather as a put: slot r of every rank's gathered buffer = rank r's rows
        lrows, kout = B * n_l, 16
        outs = [torch.randn(lrows, kout) for _ in range(P)]
        gathered = [torch.full((P, lrows, kout), float("nan")) for _ in range(P)]
        for r in range(P):
            PL.apply_peer_put_cpu(PL.peer_gather_desc(lrows, kout, r), outs[r], gathered)
        for r in range(P):
            assert torch.equal(gathered[r], torch.stack(outs, 0)), (P, r)
    with pytest.raises(AssertionError):
        PL.apply_peer_put_cpu(PL.peer_put_out_desc(1, 8, 2, 0, 16, 8), torch.zeros(4, 8).t(), [torch.zeros(4, 16)] * 2)


def test_peer_put_kernel_copy_loop_on_the_host_equals_the_descriptor_reference():
    """the index arithmetic of peer_put_kernel itself: its copy loop (csrc/peer_exchange_body.cuh, shared source) is
    compiled with g++ and run for every (destination, thread id) of an emulated launch against host buffers standing in
    for the peers; the result must equal the descriptor reference checked above (test infrastructure only)"""
    import ctypes as C
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "host_emu", "build", "libpeer_emu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-fPIC", "-shared",
                    os.path.join(here, "host_emu", "peer_emu.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.emu_peer_put.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64), C.c_int64, C.c_int64, C.c_int, C.c_int64]

    def run(desc, src, dsts, threads):
        es = src.element_size()
        ptrs = (C.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts])
        st = lib.emu_peer_put(src.data_ptr(), ptrs, len(dsts), (C.c_int * 4)(*desc["n"]),
                              (C.c_int64 * 4)(*[s * es for s in desc["ss"]]), (C.c_int64 * 4)(*[s * es for s in desc["ds"]]),
                              desc["src_peer_stride"] * es, desc["dst_off"] * es, desc["item"] * es, threads)
        assert st == 0

    for P, threads in ((2, 32), (4, 992), (8, 64)):     # whole warps: the copy is one warp per item
        torch.manual_seed(P)
        B, Fr, fs, H, d = 2, 3, 16, 8, 8
        D = H * d
        n, n_l, Dl = Fr * fs, Fr * fs // P, (H // P) * d
        for r in range(P):
            qkv = torch.randn(B * n_l, 3 * D).bfloat16()
            attf = torch.randn(B, n, Dl).bfloat16()
            for desc, src, shape in ((PL.peer_put_qkv_desc(B, n, P, r, D, Dl), qkv, (B, n, 3 * Dl)),
                                     (PL.peer_put_out_desc(B, n, P, r, D, Dl), attf, (B * n_l, D))):
                ref = [torch.zeros(shape, dtype=torch.bfloat16) for _ in range(P)]
                got = [torch.zeros(shape, dtype=torch.bfloat16) for _ in range(P)]
                PL.apply_peer_put_cpu(desc, src, ref)
                run(desc, src, got, threads)
                for q in range(P):
                    assert torch.equal(got[q], ref[q]), (P, r, q)
