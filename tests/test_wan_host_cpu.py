"""CPU dry run of the Wan launch lists at the 14B / 720p shapes of BASELINE.json configs[4] (one transformer block):
every C-ABI call validates its arguments on the host (no kernel can run here, so each call must fail with FG_ERR_CUDA,
never FG_ERR_INVALID), single GPU and as rank 1 of a 4-way Ulysses group."""
import pytest
import torch


@pytest.mark.parametrize("P,rank,with_r", [(1, 0, False), (4, 1, False), (1, 0, True)])
def test_wan14b_720p_launch_list_arguments_validate_on_the_host(monkeypatch, P, rank, with_r):
    from fastgen_b200 import _lib as L
    from fastgen_b200 import ops
    from fastgen_b200 import wan as W

    monkeypatch.setattr(W, "_ALLOW_HOST_PLANS", True)
    monkeypatch.setattr(ops, "_need_cuda", lambda *a, **k: None)
    cfg = dict(W.WAN_CONFIGS[W.Wan.MODEL_ID_14B], layers=1)
    torch.set_default_dtype(torch.bfloat16)
    try:
        net = W.Wan(transformer_config=cfg, enable_logvar_linear=False, r_timestep=with_r, encoder_depth=None).eval()
    finally:
        torch.set_default_dtype(torch.float32)
    monkeypatch.setattr(net, "_sp_size", lambda: P)
    monkeypatch.setattr(net, "_sp_rank", lambda: rank)
    net.sp_group = object() if P > 1 else None
    net.repack()
    B, F, Hh, Ww = 1, 3, 90, 160
    statuses = []

    def dry_run(plan):
        for fn, args, desc, _ in plan.steps:
            if args is None:
                continue     # a torch.distributed collective
            st = fn(*args, None)
            statuses.append((fn.__name__, desc, st, L.lib().fg_last_error().decode() if st else ""))

    monkeypatch.setattr(W._Plan, "run", lambda self, stream, profile=None: dry_run(self))
    import types
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace(cuda_stream=0))
    cross = net._cross_kv(torch.zeros(B, 512, 4096, dtype=torch.bfloat16), "__bidir__")
    plan = net._build_plan(B, F, Hh, Ww, torch.bfloat16, False, None, 0, False, 0, None, cross, has_r=with_r)
    assert plan.n_head == (11 if with_r else 7)
    dry_run(plan)
    assert len(statuses) > 20
    invalid = [s for s in statuses if s[2] == -1]
    assert not invalid, invalid[:5]
    assert all(s[2] != 0 for s in statuses), "a launch succeeded without a GPU?"
    if P > 1:
        assert any(s[1] == "a2a_pack" for s in statuses) and any(s[1] == "unshard" for s in statuses)


def test_wan14b_peer_exchange_launch_list_validates_on_the_host(monkeypatch):
    """the opt-in one-pass exchange (fg_peer_put / fg_peer_wait instead of pack + NCCL all-to-all + unpack): the launch
    list of rank 1 of a 4-way group has a put + wait pair per direction per block, no a2a pack / unpack copies and no
    collective inside the block, and every call's arguments validate (FG_ERR_CUDA here, never FG_ERR_INVALID)"""
    import types

    from fastgen_b200 import _lib as L
    from fastgen_b200 import ops
    from fastgen_b200 import wan as W

    P, rank = 4, 1
    monkeypatch.setattr(W, "_ALLOW_HOST_PLANS", True)
    monkeypatch.setattr(ops, "_need_cuda", lambda *a, **k: None)
    cfg = dict(W.WAN_CONFIGS[W.Wan.MODEL_ID_14B], layers=2)
    torch.set_default_dtype(torch.bfloat16)
    try:
        net = W.Wan(transformer_config=cfg, enable_logvar_linear=False).eval()
    finally:
        torch.set_default_dtype(torch.float32)
    monkeypatch.setattr(net, "_sp_size", lambda: P)
    monkeypatch.setattr(net, "_sp_rank", lambda: rank)
    net.sp_group = object()
    net.sp_exchange = "peer"

    def host_peer_buffers(B, n, D, Dl, P_, dev, kout):
        # stand-in for the symmetric-memory rendezvous: every "peer" is a region of host memory
        e_full, e_att, e_gat = B * n * 3 * Dl, B * (n // P_) * D, P_ * B * (n // P_) * kout
        bufs = [torch.zeros(e_full + e_att + e_gat + 96, dtype=torch.bfloat16) for _ in range(P_)]
        off_att, off_gat, off_flags = e_full * 2, (e_full + e_att) * 2, (e_full + e_att + e_gat) * 2
        ptrs = [b.data_ptr() for b in bufs]
        mine = bufs[rank]
        return dict(qkv_full=mine[:e_full].view(B, n, 3 * Dl), att=mine[e_full:e_full + e_att].view(B * (n // P_), D),
                    tok_out_g=mine[e_full + e_att:e_full + e_att + e_gat].view(P_, B * (n // P_), kout),
                    peer=dict(buf=bufs, hdl=None, full_ptrs=ptrs, att_ptrs=[p + off_att for p in ptrs],
                              gat_ptrs=[p + off_gat for p in ptrs],
                              flag_ptrs=[[p + off_flags + c * 64 for p in ptrs] for c in range(3)],
                              state=torch.zeros(3, 4, dtype=torch.int64)))

    monkeypatch.setattr(net, "_peer_buffers", host_peer_buffers)
    net.repack()
    B, F, Hh, Ww = 1, 3, 90, 160
    statuses = []

    def dry_run(plan):
        for fn, args, desc, _ in plan.steps:
            if args is None:
                statuses.append(("collective", desc, None, ""))
                continue
            st = fn(*args, None)
            statuses.append((fn.__name__, desc, st, L.lib().fg_last_error().decode() if st else ""))

    monkeypatch.setattr(W._Plan, "run", lambda self, stream, profile=None: dry_run(self))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace(cuda_stream=0))
    cross = net._cross_kv(torch.zeros(B, 512, 4096, dtype=torch.bfloat16), "__bidir__")
    plan = net._build_plan(B, F, Hh, Ww, torch.bfloat16, False, None, 0, False, 0, None, cross)
    statuses.clear()
    dry_run(plan)
    descs = [s[1] for s in statuses]
    # (put and wait are one launch: fg_peer_put with wait_flags set)
    assert descs.count("peer_put_qkv") == 2 and descs.count("peer_put_out") == 2 and "peer_wait_qkv" not in descs
    assert "a2a_pack" not in descs and "a2a_unpack" not in descs and "a2a_qkv" not in descs and "a2a_out" not in descs
    # no torch.distributed call is left anywhere in the launch list: the final gather is a put + wait as well
    assert not any(s[0] == "collective" for s in statuses)
    j = descs.index("peer_put_gather")
    assert descs[j - 1:j + 3] == ["proj_out", "peer_put_gather", "unshard", "unpatchify"]
    # put -> wait -> attention -> put -> wait -> output projection, in that order inside a block
    i = descs.index("peer_put_qkv")
    assert descs[i:i + 4] == ["peer_put_qkv", "self_attn", "peer_put_out", "o1"]
    invalid = [s for s in statuses if s[2] == -1]
    assert not invalid, invalid[:5]
    assert all(s[2] != 0 for s in statuses if s[0] != "collective"), "a launch succeeded without a GPU?"
    # misuse is rejected on the host: unknown mode, too many peers
    net.sp_exchange = "bogus"
    net._ws.clear()
    with pytest.raises(ValueError, match="sp_exchange"):
        net._workspace(B, F, Hh, Ww, torch.bfloat16)
    with pytest.raises(ValueError, match="peers"):
        ops.peer_put_params(torch.zeros(8), dict(n=[1] * 4, ss=[0] * 4, ds=[0] * 4, item=8, src_peer_stride=0, dst_off=0),
                            [1] * 9, [1] * 9, 0, torch.zeros(4, dtype=torch.int64))
