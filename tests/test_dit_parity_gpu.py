"""Parity of the CUDA DiT path against the CPU oracle (oracle/dit_oracle.py, itself pinned bit-exactly to the unmodified
reference DiT run with stand-ins for timm's three classes: tests/golden/dit_reference.npz) on identical random weights
and inputs; bf16 path tolerance rel-L2 <= 1e-2 on x0."""
import pytest
import torch

from oracle import dit_oracle as DO
from oracle import sampler_oracle as SO

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _net(cfg, seed):
    from fastgen_b200.dit import DiT

    sd = DO.make_state_dict(cfg, seed)
    kw = {k: v for k, v in cfg.items()}
    net = DiT(**kw)
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return net.cuda().eval(), sd


@pytest.mark.parametrize("cfg", [DO.tiny_config(), DO.tiny_config(hidden_size=144, num_heads=2, depth=3),
                                 DO.tiny_config(use_sit_convention=True)], ids=["d64", "d72", "sit"])
def test_dit_forward_matches_oracle(cfg):
    net, sd = _net(cfg, 0)
    g = torch.Generator().manual_seed(1)
    B, S = 3, cfg["input_size"]
    x = torch.randn(B, cfg["in_channels"], S, S, generator=g)
    t = torch.tensor([0.9, 0.5, 0.1], dtype=torch.float64)
    lab = torch.eye(cfg["num_classes"])[torch.tensor([1, 7, 3])]
    lab[2] = 0  # all-zero row -> the null class
    for pred in ("x0", "flow"):
        out = net(x.cuda(), t.cuda(), condition=lab.cuda(), fwd_pred_type=pred)
        ref = DO.dit_forward(sd, cfg, x, t, lab, fwd_pred_type=pred)
        assert out.shape == x.shape and out.dtype == x.dtype
        assert _rel(out, ref) < 1e-2, (pred, _rel(out, ref))


def test_dit_two_step_sampler_matches_oracle():
    from fastgen_b200.sampler import FastGenModel

    cfg = DO.tiny_config()
    net, sd = _net(cfg, 2)
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(2, 4, 8, 8, generator=g)
    lab = torch.eye(10)[torch.tensor([2, 5])]
    eps = torch.randn(2, 4, 8, 8, generator=g)
    sched = SO.Schedule("rf")
    ref = SO.generator_fn(lambda x, tb, **kw: DO.dit_forward(sd, cfg, x, tb, lab), sched, noise, 2, None, [eps], "sde")
    orig = torch.randn_like
    torch.randn_like = lambda x, **kw: eps.to(x.device, x.dtype)
    try:
        out = FastGenModel.generator_fn(net, noise.cuda(), condition=lab.cuda(), student_sample_steps=2)
    finally:
        torch.randn_like = orig
    assert _rel(out, ref) < 1e-2, _rel(out, ref)


@pytest.mark.parametrize("tct", ["abs", "diff"])
def test_dit_r_embedder_matches_oracle(tct):
    """second timestep r (MeanFlow-style): c = t_emb + y + r_embedder(r | t - r), reference DiT/network.py:518-533"""
    cfg = DO.tiny_config(r_timestep=True, time_cond_type=tct)
    net, sd = _net(cfg, 3)
    g = torch.Generator().manual_seed(5)
    B, S = 2, cfg["input_size"]
    x = torch.randn(B, cfg["in_channels"], S, S, generator=g)
    t = torch.tensor([0.9, 0.6], dtype=torch.float64)
    r = torch.tensor([0.2, 0.6], dtype=torch.float64)
    lab = torch.eye(cfg["num_classes"])[torch.tensor([4, 0])]
    out = net(x.cuda(), t.cuda(), condition=lab.cuda(), r=r.cuda(), fwd_pred_type="flow")
    ref = DO.dit_forward(sd, cfg, x, t, lab, fwd_pred_type="flow", r=r, time_cond_type=tct)
    assert _rel(out, ref) < 1e-2, _rel(out, ref)
    out0 = net(x.cuda(), t.cuda(), condition=lab.cuda(), fwd_pred_type="flow")
    ref0 = DO.dit_forward(sd, cfg, x, t, lab, fwd_pred_type="flow")
    assert _rel(out0, ref0) < 1e-2 and _rel(out0, ref) > 3e-2
