"""CPU dry run of the EDM launch lists (BASELINE.json configs[0] and [1] plus the tiny golden-fixture networks): every C-ABI
call validates its arguments on the host (no kernel can run here: each must fail with FG_ERR_CUDA, never FG_ERR_INVALID),
and the planner's bookkeeping is checked: every GroupNorm has exactly one statistics source, and the full-size networks
need no separate fg_gn_stats pass (the producing convs accumulate the statistics, skip concatenations included)."""
import collections

import pytest

CIFAR = dict(img_resolution=32, img_channels=3, label_dim=10, sigma_shift=0.0, sigma_data=0.5, model_type="SongUNet",
             augment_dim=9, model_channels=128, channel_mult=[2, 2, 2], channel_mult_noise=1,
             embedding_type="positional", encoder_type="standard", decoder_type="standard", resample_filter=[1, 1],
             dropout=0.0, label_dropout=0, r_timestep=False, drop_precond=None)
IN64 = dict(img_resolution=64, img_channels=3, label_dim=1000, sigma_shift=0.0, sigma_data=0.5,
            model_type="DhariwalUNet", augment_dim=0, model_channels=192, channel_mult=[1, 2, 3, 4], channel_mult_emb=4,
            num_blocks=3, attn_resolutions=[32, 16, 8], dropout=0.0, label_dropout=0, r_timestep=False,
            drop_precond=None)


def _cfgs():
    from oracle.make_golden import CONFIGS

    return [("cifar", CIFAR, "bf16", 0, "bf16"), ("in64", IN64, "bf16", 0, "bf16"), ("in64_x3", IN64, "bf16x3", 0, "bf16"),
            ("song_tiny", CONFIGS["song_tiny"], "bf16", None, "bf16"),
            ("dhariwal_tiny", CONFIGS["dhariwal_tiny"], "bf16x3", None, "bf16"),
            # fp32 CUDA-core attention on the split-precision path: head widths 256 (CIFAR), 64 (IN-64, tiny)
            ("cifar_x3_f32attn", CIFAR, "bf16x3", 0, "f32"), ("in64_x3_f32attn", IN64, "bf16x3", 0, "f32"),
            ("dhariwal_tiny_x3_f32attn", CONFIGS["dhariwal_tiny"], "bf16x3", None, "f32"),
            ("song_tiny_x3_f32attn", CONFIGS["song_tiny"], "bf16x3", None, "f32"),
            # head width 32 has no fg_attention_f32 instance: that block keeps the bf16 tensor-core kernel
            ("narrow_heads_x3", dict(CONFIGS["song_tiny"], model_channels=32, channel_mult=[1, 1]), "bf16x3", None, "f32")]


@pytest.mark.parametrize("name,cfg,compute,max_stats_passes,x3_attention", _cfgs(), ids=[c[0] for c in _cfgs()])
def test_edm_launch_list_arguments_validate_on_the_host(monkeypatch, name, cfg, compute, max_stats_passes, x3_attention):
    from fastgen_b200 import _lib as L
    from fastgen_b200 import edm as E
    from fastgen_b200 import ops

    monkeypatch.setattr(E, "_ALLOW_HOST_PLANS", True)
    monkeypatch.setattr(ops, "_need_cuda", lambda *a, **k: None)
    net = E.EDMPrecond(**cfg).eval()
    net.set_compute(compute)
    net.set_x3_attention(x3_attention)
    if name == "narrow_heads_x3":
        with pytest.warns(UserWarning, match="no fp32 attention kernel for head width 32"):
            plan = net._plan(3)
    else:
        plan = net._plan(3)
    counts = collections.Counter(fn.__name__ for fn, _ in plan.steps)
    n_attn = sum(1 for md in (net.model.enc, net.model.dec) for b in md.values() if getattr(b, "num_heads", 0))
    f32 = compute == "bf16x3" and x3_attention == "f32" and name != "narrow_heads_x3"
    assert counts["fg_attention_f32" if f32 else "fg_attention"] == n_attn > 0
    assert counts["fg_attention" if f32 else "fg_attention_f32"] == 0
    invalid, ok = [], 0
    for fn, args in plan.steps:
        st = fn(*args, None)
        if st == -1:
            invalid.append((fn.__name__, L.lib().fg_last_error().decode()))
        ok += st == 0
    assert not invalid, invalid[:5]
    assert ok <= counts["fg_zero"], "a kernel launch succeeded without a GPU?"   # (cudaMemsetAsync may report success)
    assert counts["fg_conv_gemm"] > 10 and counts["fg_gn_apply"] > 10
    if max_stats_passes is not None:
        assert counts["fg_gn_stats"] <= max_stats_passes, counts
