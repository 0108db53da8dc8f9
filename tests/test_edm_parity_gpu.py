"""Parity of the CUDA EDM path (through the C ABI) against the CPU oracle and the reference's golden fixtures.

Tolerances are the ones BASELINE.json's north star states: per-step x0 relative L2 <= 1e-2 on the bf16 path and
<= 1e-4 on the fp32 ("bf16x3" split precision) path; t lists / index tensors bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import edm_oracle as EO
from oracle import sampler_oracle as SO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
# bf16x3: convolutions with three bf16 products + fp32 CUDA-core attention (fg_attention_f32).  Measured on B200: forward
# 1.0e-5 / 1.4e-5 on the two tiny networks (gpurun_out/x3_f32_attention.txt, profiles/x3_f32_attention_r01.md); the CPU
# emulation of the rounding points predicts <= 1.6e-5 for every quantity asserted below (profiles/x3_emulation.py --golden)
TOL = {"bf16": 1e-2, "bf16x3": 1e-4}


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _net(cfg, seed, compute):
    from fastgen_b200.edm import EDMPrecond

    net = EDMPrecond(compute=compute, **cfg)
    sd = EO.make_state_dict(cfg, seed)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all("resample_filter" in k for k in missing), (missing, unexpected)
    return net.cuda().eval(), sd


class _FixedDraws:
    """make torch.randn_like return the reference's recorded draws, in order"""

    def __init__(self, eps):
        self.eps = [e.cuda() for e in eps]
        self.i = 0

    def __enter__(self):
        self._orig = torch.randn_like

        def rl(x, **kw):
            e = self.eps[self.i].to(x.dtype)
            self.i += 1
            assert e.shape == x.shape
            return e

        torch.randn_like = rl
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


@pytest.mark.parametrize("compute", ["bf16", "bf16x3"])
@pytest.mark.parametrize("name", ["song_tiny", "dhariwal_tiny"])
def test_forward_matches_reference_golden(name, compute):
    g = _load(f"edm_{name}.npz")
    cfg = json.loads(str(g["cfg"]))
    net, _ = _net(cfg, int(g["seed"]), compute)
    x_t, t, labels = _t(g["x_t"]).cuda(), _t(g["t"]).cuda(), _t(g["labels"]).cuda()
    out = net(x_t, t, condition=labels, fwd_pred_type="x0")
    assert out.dtype == x_t.dtype and out.shape == x_t.shape
    assert _rel(out, _t(g["fwd_x0"])) < TOL[compute], _rel(out, _t(g["fwd_x0"]))
    out_nl = net(x_t, t, condition=None, fwd_pred_type="x0")
    assert _rel(out_nl, _t(g["fwd_x0_nolabel"])) < TOL[compute]
    out_eps = net(x_t, t, condition=labels, fwd_pred_type="eps")
    assert _rel(out_eps, _t(g["fwd_eps"])) < 5 * TOL[compute]
    # bf16 latents in -> bf16 out (dtype rule of the reference)
    out_b = net(x_t.bfloat16(), t, condition=labels, fwd_pred_type="x0")
    assert out_b.dtype == torch.bfloat16
    assert _rel(out_b.float(), _t(g["fwd_x0"])) < 2e-2


@pytest.mark.parametrize("compute", ["bf16", "bf16x3"])
@pytest.mark.parametrize("name", ["song_tiny", "dhariwal_tiny"])
@pytest.mark.parametrize("tag,steps,stype,tl", [("gen1_sde", 1, "sde", None), ("gen2_sde", 2, "sde", None),
                                                ("gen2_ode", 2, "ode", None), ("gen2_sde_tl", 2, "sde", [80.0, 1.1, 0.0]),
                                                ("gen4_sde", 4, "sde", None)])
def test_sampler_matches_reference_golden(name, compute, tag, steps, stype, tl):
    from fastgen_b200.sampler import FastGenModel

    g = _load(f"edm_{name}.npz")
    cfg = json.loads(str(g["cfg"]))
    net, _ = _net(cfg, int(g["seed"]), compute)
    noise, labels = _t(g["noise"]).cuda(), _t(g["labels"]).cuda()
    eps = [_t(g[f"{tag}_eps{i}"]) for i in range(steps) if f"{tag}_eps{i}" in g]
    with _FixedDraws(eps) as d:
        out = FastGenModel.generator_fn(net, noise, condition=labels, student_sample_steps=steps,
                                        student_sample_type=stype, t_list=tl)
    assert d.i == len(eps)
    assert out.dtype == noise.dtype
    # multi-step error compounds through the re-noised x_t: allow 2x the single-step budget at the end
    assert _rel(out, _t(g[f"{tag}_out"])) < 2 * TOL[compute], _rel(out, _t(g[f"{tag}_out"]))


def test_reference_generator_contract_drives_our_network():
    """the (oracle restatement of the) reference loop only needs net(x,t,...) + noise_scheduler methods"""
    g = _load("edm_song_tiny.npz")
    cfg = json.loads(str(g["cfg"]))
    net, _ = _net(cfg, int(g["seed"]), "bf16")
    noise, labels = _t(g["noise"]).cuda(), _t(g["labels"]).cuda()
    sched = net.noise_scheduler
    tl = sched.get_t_list(2, device=noise.device)
    assert torch.equal(tl.cpu(), SO.Schedule("edm").get_t_list(2))      # bit-exact t list
    x = sched.latents(noise, tl[0])
    assert torch.equal(x.cpu(), SO.Schedule("edm").latents(noise.cpu(), tl[0].cpu()))
    x0 = net(x, tl[0].expand(4), condition=labels, fwd_pred_type="x0")
    e = _t(g["gen2_sde_eps0"]).cuda()
    xn = sched.forward_process(x0, e, tl[1].expand(4))
    assert torch.equal(xn.cpu(), SO.Schedule("edm").forward_process(x0.cpu(), e.cpu(), tl[1].expand(4).cpu()))
    eps2 = sched.x0_to_eps(xt=x, x0=x0, t=tl[0].expand(4))
    assert torch.equal(eps2.cpu(), SO.Schedule("edm").x0_to_eps(x.cpu(), x0.cpu(), tl[0].expand(4).cpu()))
    with pytest.raises(AssertionError):
        sched.forward_process(x0, e, torch.full((4,), 81.0, device="cuda", dtype=torch.float64))
    with pytest.raises(AssertionError):
        net(x, tl[0].expand(4), condition=labels, fwd_pred_type="bogus")


def test_state_dict_keys_match_reference():
    from fastgen_b200.edm import EDMPrecond

    for name in ("song_tiny", "dhariwal_tiny", "meanflow_tiny"):
        g = _load(f"edm_{name}.npz")
        cfg = json.loads(str(g["cfg"]))
        net = EDMPrecond(**cfg)
        ours = {k: tuple(v.shape) for k, v in net.state_dict().items() if "resample_filter" not in k}
        assert ours == EO.param_shapes(cfg)


@pytest.mark.parametrize("tag,steps,stype", [("gen1_sde", 1, "sde"), ("gen2_sde", 2, "sde"), ("gen2_ode", 2, "ode"),
                                             ("gen4_ode", 4, "ode")])
def test_meanflow_sampler_matches_reference_golden(tag, steps, stype):
    from fastgen_b200.sampler import MeanFlowModel

    g = _load("edm_meanflow_tiny.npz")
    cfg = json.loads(str(g["cfg"]))
    net, _ = _net(cfg, int(g["seed"]), "bf16")
    noise, labels = _t(g["noise"]).cuda(), _t(g["labels"]).cuda()
    flow = net(noise, _t(g["t"]).cuda(), condition=labels, r=_t(g["r"]).cuda(), fwd_pred_type="flow")
    assert _rel(flow, _t(g["fwd_flow"])) < 1e-2
    x0 = net(noise, _t(g["t"]).cuda(), condition=labels, r=_t(g["r"]).cuda(), fwd_pred_type="x0")
    assert _rel(x0, _t(g["fwd_x0"])) < 1e-2
    eps = [_t(g[f"{tag}_eps{i}"]) for i in range(steps) if f"{tag}_eps{i}" in g]
    with _FixedDraws(eps):
        out = MeanFlowModel.generator_fn(net, noise, condition=labels, student_sample_steps=steps,
                                         student_sample_type=stype)
    assert _rel(out, _t(g[f"{tag}_out"])) < 2e-2


CIFAR = dict(img_resolution=32, img_channels=3, label_dim=10, sigma_shift=0.0, sigma_data=0.5, model_type="SongUNet",
             augment_dim=9, model_channels=128, channel_mult=[2, 2, 2], channel_mult_noise=1,
             embedding_type="positional", encoder_type="standard", decoder_type="standard", resample_filter=[1, 1],
             dropout=0.0, label_dropout=0, r_timestep=False, drop_precond=None)
IN64 = dict(img_resolution=64, img_channels=3, label_dim=1000, sigma_shift=0.0, sigma_data=0.5,
            model_type="DhariwalUNet", augment_dim=0, model_channels=192, channel_mult=[1, 2, 3, 4], channel_mult_emb=4,
            num_blocks=3, attn_resolutions=[32, 16, 8], dropout=0.0, label_dropout=0, r_timestep=False,
            drop_precond=None)


@pytest.mark.parametrize("cfg,batch,steps,tl", [(CIFAR, 4, 1, None), (IN64, 2, 2, [80.0, 1.1, 0.0])],
                         ids=["cifar10_dmd2_1step", "imagenet64_scm_2step"])
def test_full_size_network_matches_oracle(cfg, batch, steps, tl):
    """BASELINE configs[0] / configs[1] networks at full width and depth, small batch, per-step x0 vs the oracle"""
    from fastgen_b200.sampler import FastGenModel

    torch.manual_seed(1)
    net, sd = _net(cfg, 0, "bf16")
    R, L = cfg["img_resolution"], cfg["label_dim"]
    noise = torch.randn(batch, 3, R, R)
    labels = torch.eye(L)[torch.randint(L, (batch,))]
    eps = [torch.randn(batch, 3, R, R) for _ in range(steps - 1)]
    trace = []
    with torch.no_grad():
        ref = SO.generator_fn(lambda x, t, **kw: EO.edm_precond_forward(sd, cfg, x, t, labels), SO.Schedule("edm"),
                              noise, steps, tl, eps, "sde", trace)
    ref_x0 = [v for k, v in trace if k == "x0"]
    # per-step x0 on the oracle's own x_t (isolates one network evaluation per step)
    tvals = tl or SO.Schedule("edm").get_t_list(steps).tolist()
    ref_xt = [SO.Schedule("edm").latents(noise, torch.tensor(tvals[0], dtype=torch.float64))] + \
        [v for k, v in trace if k == "x_next"]
    for i, (xt, x0r) in enumerate(zip(ref_xt, ref_x0)):
        t = torch.full((batch,), tvals[i], dtype=torch.float64).cuda()
        x0 = net(xt.cuda(), t, condition=labels.cuda(), fwd_pred_type="x0")
        assert _rel(x0, x0r) < 1e-2, (i, _rel(x0, x0r))
    with _FixedDraws(eps):
        out = FastGenModel.generator_fn(net, noise.cuda(), condition=labels.cuda(), student_sample_steps=steps,
                                        t_list=tl)
    assert _rel(out, ref) < 2e-2, _rel(out, ref)


def test_batch_independence_at_bench_batch():
    """size-independent property at a bench-sized batch (the oracle would need minutes there): sample i of a large
    batch equals the same sample run in a batch of 2 -- no cross-sample coupling, no tile-boundary effects -- and the
    network is a deterministic function of its inputs.  Measured on B200 (profiles/determinism_probe.py, 8 repeats, bf16
    and bf16x3): BIT-IDENTICAL in both respects.  GroupNorm partial sums are merged with fp64 atomics whose order varies,
    i.e. at the 1e-16 level before the statistics are rounded to fp32; the bound leaves room for one such flip
    (a flipped bf16 rounding of one activation: ~1e-7 of the output) instead of asserting bit equality."""
    torch.manual_seed(2)
    net, _ = _net(IN64, 0, "bf16")
    B = 64
    x = torch.randn(B, 3, 64, 64, device="cuda") * 80
    t = torch.full((B,), 80.0, device="cuda", dtype=torch.float64)
    lab = torch.eye(1000, device="cuda")[torch.randint(1000, (B,), device="cuda")]
    big = net(x, t, condition=lab, fwd_pred_type="x0")
    assert torch.isfinite(big).all()
    for lo in (0, 10, 62):
        small = net(x[lo:lo + 2].clone(), t[:2], condition=lab[lo:lo + 2], fwd_pred_type="x0")
        assert _rel(big[lo:lo + 2], small) < 1e-5
    # and the network is a deterministic function of its inputs run to run
    again = net(x, t, condition=lab, fwd_pred_type="x0")
    assert _rel(again, big) < 1e-5


@pytest.mark.parametrize("compute,tol", [("bf16", 1e-2), ("bf16x3", 2e-3)])
def test_reference_checkpoint_drives_our_network(compute, tol):
    """a `.pth` in the reference Checkpointer's layout (written from a real reference EDMPrecond,
    oracle/make_golden.py:checkpoint_fixture) loaded by fastgen_b200.checkpoint reproduces the reference's forward"""
    import json
    import os

    import numpy as np

    from fastgen_b200.checkpoint import load_checkpoint
    from fastgen_b200.edm import EDMPrecond

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    io = np.load(os.path.join(gold, "ref_checkpoint_tiny_io.npz"))
    net = EDMPrecond(**json.loads(str(io["cfg"])))
    assert load_checkpoint({"net": net}, os.path.join(gold, "ref_checkpoint_tiny.pth")) == 7
    net = net.cuda().eval()
    net.set_compute(compute)
    out = net(torch.from_numpy(io["x"]).cuda(), torch.from_numpy(io["t"]).cuda(),
              condition=torch.from_numpy(io["lab"]).cuda(), fwd_pred_type="x0")
    assert _rel(out, torch.from_numpy(io["out"])) < tol, _rel(out, torch.from_numpy(io["out"]))


@pytest.mark.parametrize("compute,tol", [("bf16", 1e-2), ("bf16x3", 2e-3)])
def test_teacher_sample_with_cfg_matches_reference(compute, tol):
    """EDMPrecond.sample (EDM/network.py:976-1026) -- 6 Euler steps with classifier-free guidance 2.0
    (batched [negative | positive] forward) and without guidance, against samples of the real reference network"""
    import json
    import os

    import numpy as np

    from fastgen_b200.checkpoint import load_checkpoint
    from fastgen_b200.edm import EDMPrecond

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    io = np.load(os.path.join(gold, "ref_checkpoint_tiny_io.npz"))
    net = EDMPrecond(**json.loads(str(io["cfg"])))
    load_checkpoint({"net": net}, os.path.join(gold, "ref_checkpoint_tiny.pth"))
    net = net.cuda().eval()
    net.set_compute(compute)
    noise, lab, neg = (torch.from_numpy(io[k]).cuda() for k in ("noise", "lab", "neg"))
    out = net.sample(noise, condition=lab, neg_condition=neg, guidance_scale=2.0, num_steps=6)
    assert out.shape == noise.shape and _rel(out, torch.from_numpy(io["sample_cfg"])) < tol
    out = net.sample(noise, condition=lab, neg_condition=None, guidance_scale=None, num_steps=6)
    assert _rel(out, torch.from_numpy(io["sample"])) < tol


def test_cuda_graph_replay_matches_eager_launches():
    """opt-in EDMPrecond.use_cuda_graphs: from its second use a batch size's launch list replays as one CUDA graph; the
    outputs equal the eager path's (to the fp64-atomic noise floor) for new inputs, new t and new labels"""
    net, _ = _net(IN64, 0, "bf16")
    torch.manual_seed(5)
    xs = [torch.randn(2, 3, 64, 64, device="cuda") * s for s in (80.0, 1.1, 5.0)]
    ts = [torch.full((2,), v, device="cuda", dtype=torch.float64) for v in (80.0, 1.1, 5.0)]
    labs = [torch.eye(1000, device="cuda")[torch.randint(1000, (2,), device="cuda")] for _ in range(3)]
    eager = [net(x, t, condition=lab, fwd_pred_type="x0").clone() for x, t, lab in zip(xs, ts, labs)]
    net.use_cuda_graphs = True
    net._plans.clear()
    replay = [net(x, t, condition=lab, fwd_pred_type="x0").clone() for x, t, lab in zip(xs, ts, labs)]
    assert net.graph_launches > 0 and net._plans[2].graph is not None
    for a, b in zip(eager, replay):
        assert _rel(b, a) < 1e-2, _rel(b, a)   # same bound as run-to-run determinism (fp64-atomic noise floor) above
