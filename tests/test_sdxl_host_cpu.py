"""CPU tests of the SDXL host side: diffusers state-dict key names, the published parameter count of the SDXL-base UNet,
scheduler constants, and a dry run of the complete launch list in which every C-ABI call validates its arguments on the
host (no kernel can run here: each call must fail with FG_ERR_CUDA, never FG_ERR_INVALID)."""
import types

import pytest
import torch

from oracle import sdxl_oracle as XO


def test_state_dict_keys_follow_diffusers_names():
    from fastgen_b200.sdxl import StableDiffusionXL, layout

    cfg = XO.tiny_config()
    sd = XO.make_state_dict(cfg, 0)
    net = StableDiffusionXL(unet_config=cfg)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    extra = sorted(set(mine) - set(sd))
    assert extra == ["unet.logvar_linear.bias", "unet.logvar_linear.weight"]     # reference SDXL/network.py:231
    assert all(mine[k] == tuple(v.shape) for k, v in sd.items())
    assert "unet.down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_out.0.weight" in mine
    assert "unet.up_blocks.0.attentions.2.transformer_blocks.0.ff.net.0.proj.bias" in mine
    assert layout(cfg) == XO.layout(cfg)
    assert net.net_pred_type == "eps" and net.schedule_type == "sdxl"
    assert net.add_time_ids.tolist() == [[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]]


def test_sdxl_base_parameter_count_is_the_published_one():
    """known answer: diffusers' SDXL-base UNet2DConditionModel has 2,567,463,684 parameters"""
    from fastgen_b200.sdxl import StableDiffusionXL

    with torch.device("meta"):
        net = StableDiffusionXL()
    n = sum(p.numel() for name, p in net.unet.named_parameters() if not name.startswith("logvar_linear"))
    assert n == 2_567_463_684


def test_sd15_parameter_count_and_keys():
    """known answer: the SD 1.5 UNet2DConditionModel has 859,520,964 parameters; 1x1-conv projections, no add_embedding;
    the oracle's state-dict factory produces exactly the module tree's keys"""
    from fastgen_b200.sdxl import SD15_BASE, StableDiffusion15

    with torch.device("meta"):
        net = StableDiffusion15()
    n = sum(p.numel() for name, p in net.unet.named_parameters() if not name.startswith("logvar_linear"))
    assert n == 859_520_964
    sd = net.state_dict()
    assert sd["unet.down_blocks.0.attentions.0.proj_in.weight"].shape == (320, 320, 1, 1)
    assert not any("add_embedding" in k for k in sd) and net.schedule_type == "sd" and net.net_pred_type == "eps"
    assert {k: tuple(v) for k, v in SD15_BASE.items() if isinstance(v, tuple)} == \
        {k: tuple(v) for k, v in XO.SD15_BASE.items() if isinstance(v, tuple)}
    tiny = XO.tiny_sd15_config()
    small = StableDiffusion15(unet_config=tiny)
    want = set(XO.make_state_dict(tiny, 0).keys())
    have = {k for k in small.state_dict().keys() if "logvar_linear" not in k}
    assert want == have


def test_scheduler_constants():
    from fastgen_b200.noise_schedule import get_noise_schedule

    abar = XO.alphas_cumprod()
    assert abs(abar[0].item() - (1 - 0.00085)) < 1e-7 and abs(abar[-1].item() - 0.0046604) < 1e-5
    ns = get_noise_schedule("sdxl")
    t = torch.tensor([0.999, 0.5, 0.0], dtype=torch.float64)
    idx = (1000 * t).long().clamp(0, 999)
    assert torch.equal(ns.alpha(t).flatten(), abar.double()[idx].sqrt())
    assert torch.equal(ns.get_t_list(4), torch.linspace(0.999, 0, 5, dtype=torch.float64))
    assert torch.equal(ns.rescale_t(t), 1000 * t)


def test_timestep_sinusoid_matches_oracle():
    from fastgen_b200.sdxl import timestep_sinusoid

    t = torch.tensor([999.0, 499.5, 0.0], dtype=torch.float64)
    assert torch.equal(timestep_sinusoid(t, 320), XO.timesteps_proj(t, 320))


@pytest.mark.parametrize("cfg,B,S", [(XO.tiny_config(), 2, 32), (XO.tiny_config(transformer_layers=(1, 2, 3)), 3, 16)])
def test_launch_list_arguments_validate_on_the_host(monkeypatch, cfg, B, S):
    from fastgen_b200 import _lib as L
    from fastgen_b200 import ops
    from fastgen_b200 import sdxl as S_

    monkeypatch.setattr(S_, "_ALLOW_HOST_PLANS", True)
    monkeypatch.setattr(ops, "_need_cuda", lambda *a, **k: None)
    net = S_.StableDiffusionXL(unet_config=cfg).eval()
    net.repack()
    statuses = []

    def dry_run(self, stream, profile=None):
        for fn, args, desc, _ in self.steps:
            st = fn(*args, None)
            statuses.append((fn.__name__, desc, st, L.lib().fg_last_error().decode() if st else ""))

    monkeypatch.setattr(S_._Plan, "run", dry_run)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace(cuda_stream=0))
    ctx = net._context(torch.randn(B, 77, cfg["cross_attention_dim"]))
    plan, ws = net._plan(B, S, S, torch.bfloat16, ctx)
    plan.run(0)
    n_tr = sum(len(a.transformer_blocks) for _, a in net._transformers())
    assert sum(1 for s in statuses if s[1] == "kv2") == n_tr
    assert sum(1 for s in statuses if s[1] == "cross_attn") == n_tr
    invalid = [s for s in statuses if s[2] == -1]     # FG_ERR_INVALID
    assert not invalid, invalid[:5]
    assert all(s[2] != 0 for s in statuses), "a launch succeeded without a GPU?"
    # every GroupNorm gets its statistics from the producing GEMM(s) -- skip concatenations included (two producers, one
    # slot) -- except the two norms per downsampler that read a subsampled tensor (next resnet + its concat consumer)
    n_gn_stats = sum(1 for s in statuses if s[0] == "fg_gn_stats")
    n_down = sum(1 for d in S_.layout(cfg)["down"] if d["down"])
    if S >= 32:   # below 32 pixels per image the producer cannot fuse the statistics (fg_conv_gemm contract)
        assert n_gn_stats <= 2 * n_down, n_gn_stats
