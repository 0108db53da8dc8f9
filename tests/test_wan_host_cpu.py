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
