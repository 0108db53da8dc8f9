"""Ulysses sequence-parallel Wan on 2 GPUs (NCCL all-to-all) against the single-GPU path: same weights, same inputs,
KV-cached chunks.  Needs two CUDA devices (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wan_oracle as WO

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _run_chunks(net, x, text, total, dev):
    outs = []
    net.clear_caches()
    for (s, e) in [(0, 4), (4, 7)]:
        xc = x[:, :, s:e].contiguous().to(dev)
        for tval, store in [(0.937, False), (0.0, True)]:
            t = torch.full((x.shape[0],), tval, dtype=torch.float64, device=dev)
            outs.append(net(xc, t, condition=text.to(dev), fwd_pred_type="x0", cur_start_frame=s, store_kv=store,
                            is_ar=True).float().cpu())
    return outs


def _worker(rank, world, port, ret, exchange="nccl"):
    import fastgen_b200.wan as W

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg = WO.tiny_config(layers=2)
        total = 7
        sd = {k: v.bfloat16() for k, v in WO.make_state_dict(cfg, 1).items()}
        g = torch.Generator().manual_seed(3)
        x = torch.randn(2, 16, total, 8, 8, generator=g).bfloat16()
        text = torch.randn(2, 24, cfg["text_dim"], generator=g).bfloat16()

        def build():
            net = W.CausalWan(transformer_config=cfg, enable_logvar_linear=False, total_num_frames=total, chunk_size=3)
            net.load_state_dict(sd, strict=True)
            return net.to(torch.bfloat16).to(dev).eval()

        single = _run_chunks(build(), x, text, total, dev)
        net = build()
        net.sp_group = dist.group.WORLD
        net.sp_exchange = "peer" if exchange.startswith("peer") else exchange
        net.peer_timeout_s = 5.0
        sp = _run_chunks(net, x, text, total, dev)
        if exchange == "peer+graphs":
            # no NCCL call is left in the launch lists: capture them (second use) and replay (third use)
            net.sp_graphs = True
            _run_chunks(net, x, text, total, dev)
            before = net.graph_launches
            sp = _run_chunks(net, x, text, total, dev)
            assert net.graph_launches > before
        if exchange.startswith("peer"):
            net.check_peer_exchange()      # raises if a device-side wait ever timed out
            descs = [s[2] for p in net._plans.values() for s in p.steps]
            assert "peer_put_qkv" in descs and "a2a_qkv" not in descs
        errs = [_rel(a, b) for a, b in zip(sp, single)]
        # head-sharded cache: this rank holds heads [rank*H/P, (rank+1)*H/P)
        assert net._kv["pos"]["k"].shape[-1] == cfg["dim"] // world
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sequence_parallel_matches_single_gpu():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        for r in range(world):
            errs = ret[r]
            assert len(errs) == 4 and max(errs) < 5e-3, (r, errs)





@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_sequence_parallel_peer_exchange_matches_single_gpu():
    """the one-pass exchange (fg_peer_put / fg_peer_wait over symmetric memory, no NCCL in the block) gives the same
    chunks as the single-GPU path (green on 2 and 4 B200s, round 2: profiles/r2_logs/)."""
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret, "peer"), nprocs=world, join=True)
        for r in range(world):
            errs = ret[r]
            assert len(errs) == 4 and max(errs) < 5e-3, (r, errs)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_sequence_parallel_peer_exchange_replays_as_cuda_graphs():
    """with the peer exchange a forward holds no torch.distributed call, so its launch list is captured into a CUDA graph
    under sequence parallelism too (green on 2 and 4 B200s, round 2)."""
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret, "peer+graphs"), nprocs=world, join=True)
        for r in range(world):
            errs = ret[r]
            assert len(errs) == 4 and max(errs) < 5e-3, (r, errs)


def _peer_kernel_worker(rank, world, port, ret):
    """fg_peer_put / fg_peer_wait alone: every rank stores a rank-specific pattern into slot `rank` of every rank's buffer
    (the all-gather descriptor), three rounds on one channel, and checks what arrived after the device-side wait."""
    import ctypes as C

    import torch.distributed._symmetric_memory as symm

    from fastgen_b200 import _lib as L
    from fastgen_b200 import ops
    from fastgen_b200 import parallel as PL

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rows, cols = 300, 64
        data_b = world * rows * cols * 2
        enable = getattr(symm, "enable_symm_mem_for_group", None)
        if enable is not None:
            enable(dist.group.WORLD.group_name)
        buf = symm.empty(data_b + 256, dtype=torch.uint8, device=dev)
        buf.zero_()
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        torch.cuda.synchronize()
        dist.barrier()
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        gathered = buf[:data_b].view(torch.bfloat16).view(world, rows, cols)
        state = torch.zeros(4, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        ok = True
        for rnd in range(3):
            src = torch.full((rows, cols), float(10 * rnd + rank + 1), device=dev, dtype=torch.bfloat16)
            p = ops.peer_put_params(src, PL.peer_gather_desc(rows, cols, rank), ptrs, [q + data_b for q in ptrs], rank, state)
            L.check(L.lib().fg_peer_put(C.byref(p), stream), "fg_peer_put")
            L.check(L.lib().fg_peer_wait(ptrs[rank] + data_b, world, state.data_ptr(), int(5e9), stream), "fg_peer_wait")
            torch.cuda.synchronize()
            want = torch.stack([torch.full((rows, cols), float(10 * rnd + r + 1)) for r in range(world)]).bfloat16()
            ok = ok and torch.equal(gathered.cpu(), want)
            dist.barrier()      # (test only) nobody overwrites a slot before everybody has checked it
        ret[rank] = (ok, state.cpu().tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(300)
def test_peer_put_and_wait_kernels_alone():
    """run this one FIRST when bringing the peer exchange up on hardware: it isolates the two kernels and the symmetric-
    memory rendezvous from the Wan plan"""
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_peer_kernel_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        for r in range(world):
            ok, state = ret[r]
            assert ok and state == [3, 0, 3, 0], (r, ok, state)     # 3 puts, counter reset, 3 waits, no timeout
