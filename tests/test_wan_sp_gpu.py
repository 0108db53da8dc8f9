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


_UNVERIFIED = pytest.mark.skipif(os.environ.get("FASTGEN_B200_TEST_UNVERIFIED") != "1",
                                 reason="never run on hardware yet: set FASTGEN_B200_TEST_UNVERIFIED=1")


@_UNVERIFIED
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_sequence_parallel_peer_exchange_matches_single_gpu():
    """the one-pass exchange (fg_peer_put / fg_peer_wait over symmetric memory, no NCCL in the block) gives the same
    chunks as the single-GPU path.  NOT yet run on hardware (written after round 1's GPU budget ended): first thing to
    run in round 2 with `gpurun --gpus 2`, under a short timeout."""
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret, "peer"), nprocs=world, join=True)
        for r in range(world):
            errs = ret[r]
            assert len(errs) == 4 and max(errs) < 5e-3, (r, errs)


@_UNVERIFIED
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_sequence_parallel_peer_exchange_replays_as_cuda_graphs():
    """with the peer exchange a forward holds no torch.distributed call, so its launch list is captured into a CUDA graph
    under sequence parallelism too.  NOT yet run on hardware; run after the test above is green."""
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret, "peer+graphs"), nprocs=world, join=True)
        for r in range(world):
            errs = ret[r]
            assert len(errs) == 4 and max(errs) < 5e-3, (r, errs)
