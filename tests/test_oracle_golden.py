"""Pin the CPU oracle (oracle/) against fixtures produced by the real reference (oracle/make_golden.py).

Runs on CPU (`-m "not gpu"`).  Integer / index tensors and everything the reference evaluates in float64 must be
bit-exact; fp32 network forwards agree to op-ordering noise."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import edm_oracle as EO
from oracle import sampler_oracle as SO

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _t(a, bf16=False):
    t = torch.from_numpy(np.asarray(a))
    return t.view(torch.bfloat16) if bf16 else t


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_edm_t_list_and_indices_bit_exact():
    g = _load("schedules.npz")
    s = SO.Schedule("edm")
    assert np.array_equal(s.sigmas, g["edm_sigmas"])
    for k in (1, 2, 3, 4, 7, 8):
        assert np.array_equal(s.t_list_indices(k), g[f"edm_idx_{k}"])
        assert np.array_equal(s.get_t_list(k).numpy(), g[f"edm_t_list_{k}"])
        assert np.array_equal(SO.Schedule("rf").get_t_list(k).numpy(), g[f"rf_t_list_{k}"])
    # SURVEY.md A.2 known answers
    assert s.t_list_indices(4).tolist() == [998, 749, 500, 251, 2]
    assert s.get_t_list(4).tolist() == [79.56382521581216, 17.49810610061577, 2.5265068484381206,
                                        0.17256643947993575, 0.0]


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("kind", ["edm", "rf", "alphas"])
def test_schedule_arithmetic_bit_exact(kind, dt):
    g = _load("schedules.npz")
    bf = dt == "bf16"
    s = SO.Schedule(kind, alphas_cumprod=_t(g["alphas_cumprod"]) if kind == "alphas" else None)
    x, e, t = _t(g[f"x_{dt}"], bf), _t(g[f"e_{dt}"], bf), _t(g[f"{kind}_t3"])
    assert torch.equal(s.latents(x, t[1]), _t(g[f"{kind}_latents_{dt}"], bf))
    assert torch.equal(s.forward_process(x, e, t), _t(g[f"{kind}_fwd_{dt}"], bf))
    assert torch.equal(s.x0_to_eps(x, e, t), _t(g[f"{kind}_x0_to_eps_{dt}"], bf))
    assert torch.equal(s.eps_to_x0(x, e, t), _t(g[f"{kind}_eps_to_x0_{dt}"], bf))
    if kind != "alphas":
        assert torch.equal(s.flow_to_x0(x, e, t), _t(g[f"{kind}_flow_to_x0_{dt}"], bf))
        assert torch.equal(s.x0_to_flow(x, e, t), _t(g[f"{kind}_x0_to_flow_{dt}"], bf))


def test_alphas_table_lookup_bit_exact():
    g = _load("schedules.npz")
    s = SO.Schedule("alphas", alphas_cumprod=_t(g["alphas_cumprod"]))
    t = _t(g["alphas_t"])
    assert np.array_equal((1000 * t).long().clamp(0, 999).numpy(), g["alphas_idx"])
    assert np.array_equal(s.alpha(t).numpy(), g["alphas_alpha"])
    assert np.array_equal(s.sigma(t).numpy(), g["alphas_sigma"])


def test_per_frame_forward_process_bit_exact():
    g = _load("schedules.npz")
    s = SO.Schedule("rf")
    out = s.forward_process(_t(g["video_x"]), _t(g["video_e"]), _t(g["video_t"])[:, None, :, None, None])
    assert torch.equal(out, _t(g["video_fwd"]))


@pytest.mark.parametrize("name", ["song_tiny", "dhariwal_tiny"])
def test_edm_forward_matches_reference(name):
    g = _load(f"edm_{name}.npz")
    cfg = json.loads(str(g["cfg"]))
    sd = EO.make_state_dict(cfg, int(g["seed"]))
    x_t, t, labels = _t(g["x_t"]), _t(g["t"]), _t(g["labels"])
    with torch.no_grad():
        out = EO.edm_precond_forward(sd, cfg, x_t, t, labels)
        out_nl = EO.edm_precond_forward(sd, cfg, x_t, t, None)
        sd64 = {k: v.double() for k, v in sd.items()}
        out64 = EO.edm_precond_forward(sd64, cfg, x_t.double(), t, labels.double())
    assert _rel(out, _t(g["fwd_x0"])) < 2e-6
    assert _rel(out_nl, _t(g["fwd_x0_nolabel"])) < 2e-6
    assert _rel(out64, _t(g["fwd_x0_f64"])) < 1e-12
    # the raw network output is NOT degenerate with the parity weights (SURVEY.md 0.6 trap)
    raw = EO.edm_precond_forward(sd, cfg, x_t, t, labels, return_raw=True)
    assert raw.std().item() > 0.1


@pytest.mark.parametrize("name", ["song_tiny", "dhariwal_tiny"])
@pytest.mark.parametrize("tag,steps,stype,tl", [("gen1_sde", 1, "sde", None), ("gen2_sde", 2, "sde", None),
                                                ("gen2_ode", 2, "ode", None), ("gen2_sde_tl", 2, "sde", [80.0, 1.1, 0.0]),
                                                ("gen4_sde", 4, "sde", None)])
def test_edm_sampler_matches_reference(name, tag, steps, stype, tl):
    g = _load(f"edm_{name}.npz")
    cfg = json.loads(str(g["cfg"]))
    sd = EO.make_state_dict(cfg, int(g["seed"]))
    noise, labels = _t(g["noise"]), _t(g["labels"])
    eps = [_t(g[f"{tag}_eps{i}"]) for i in range(steps) if f"{tag}_eps{i}" in g]
    trace = []
    with torch.no_grad():
        out = SO.generator_fn(lambda x, t, **kw: EO.edm_precond_forward(sd, cfg, x, t, labels), SO.Schedule("edm"),
                              noise, steps, tl, eps, stype, trace)
    x0s = [v for k, v in trace if k == "x0"]
    for i, v in enumerate(x0s):
        assert _rel(v, _t(g[f"{tag}_x0_{i}"])) < 1e-5, (i, _rel(v, _t(g[f"{tag}_x0_{i}"])))
    assert _rel(out, _t(g[f"{tag}_out"])) < 1e-5


@pytest.mark.parametrize("tag,steps,stype", [("gen1_sde", 1, "sde"), ("gen2_sde", 2, "sde"), ("gen2_ode", 2, "ode"),
                                             ("gen4_ode", 4, "ode")])
def test_meanflow_sampler_matches_reference(tag, steps, stype):
    g = _load("edm_meanflow_tiny.npz")
    cfg = json.loads(str(g["cfg"]))
    sd = EO.make_state_dict(cfg, int(g["seed"]))
    noise, labels = _t(g["noise"]), _t(g["labels"])
    sched = SO.Schedule("rf")
    with torch.no_grad():
        flow = EO.edm_precond_forward(sd, cfg, noise, _t(g["t"]), labels, r=_t(g["r"]))
        assert _rel(flow, _t(g["fwd_flow"])) < 2e-6
        eps = [_t(g[f"{tag}_eps{i}"]) for i in range(steps) if f"{tag}_eps{i}" in g]
        tl = sched.get_t_list(steps)
        lat = sched.latents(noise, tl[0])
        out = SO.meanflow_sample_loop(lambda x, t, r: EO.edm_precond_forward(sd, cfg, x, t, labels, r=r), sched, lat, tl,
                                      eps, stype)
    assert _rel(out, _t(g[f"{tag}_out"])) < 1e-5


def test_causal_chunking_matches_reference_trace():
    g = _load("causal_trace.npz")
    trace = json.loads(str(g["trace"]))
    calls = [c for c in trace if c[0] == "net"]
    chunks = SO.causal_chunks(7, 3)
    assert chunks == [(0, 4), (4, 7)]
    # per chunk: 4 denoise calls (store_kv False) + 1 cache-fill call at t=0 (store_kv True)
    assert trace[0] == ["clear_caches"] and trace[-1] == ["clear_caches"]
    exp_t = [0.999, 0.937, 0.833, 0.624, 0.0]
    for ci, (s, e) in enumerate(chunks):
        for k in range(5):
            c = calls[ci * 5 + k]
            assert c[1][2] == e - s and c[5] == s
            assert abs(c[2] - exp_t[k]) < 1e-12
            assert c[6] == (k == 4) and c[7] is True and c[3] == "x0" and c[4] == "pos"
            assert c[8] == ["neg_condition"]


def test_extrapolation_matches_reference():
    """generator_fn_extrapolation (causvid.py:187-397): the oracle restatement reproduces the reference's output and
    call trace (stub network + stub VAE, recorded random draws) for the sde / overlap and the ode / no-overlap runs"""
    g = _load("extrapolation_trace.npz")
    noise = torch.from_numpy(g["noise"])
    sched = SO.Schedule("rf")
    calls = []

    def net(x, t, cur_start_frame=0, store_kv=False):
        calls.append(["net", list(x.shape), float(t[0]), "x0", "pos", int(cur_start_frame), bool(store_kv), True, []])
        return 0.5 * x + 0.25 * float(t[0])

    def dec(z):
        calls.append(["vae.decode", list(z.shape)])
        return 2.0 * z

    def enc(v):
        calls.append(["vae.encode", list(v.shape)])
        return 0.5 * v + 0.125

    draws = [torch.from_numpy(g[f"eps{i}"]) for i in range(10)]
    out = SO.causvid_extrapolation(net, sched, noise, [0.999, 0.6, 0.0], 3, 3, 3, draws, "sde", 0.1, dec, enc,
                                   clear_caches=lambda: calls.append(["clear_caches"]))
    assert torch.equal(out, torch.from_numpy(g["out"]))
    assert calls == json.loads(str(g["trace"]))
    out2 = SO.causvid_extrapolation(net, sched, noise, [0.999, 0.6, 0.0], 3, 2, 0, [torch.from_numpy(g["ode_eps0"])], "ode")
    assert torch.equal(out2, torch.from_numpy(g["out_ode"]))


def test_teacher_sampler_matches_reference():
    """EDMPrecond.sample (EDM/network.py:976-1026, Euler + classifier-free guidance) on the reference network stored in
    the checkpoint fixture: the restated network + restated sampler reproduce the reference's samples"""
    import json
    import os

    io = _load("ref_checkpoint_tiny_io.npz")
    cfg = dict(sigma_shift=0.0, sigma_data=0.5, r_timestep=False, drop_precond=None, channel_mult_emb=4)
    cfg.update(json.loads(str(io["cfg"])))
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sd = torch.load(os.path.join(gold, "ref_checkpoint_tiny.pth"), weights_only=False)["model"]["net"]
    lab, noise, neg = (torch.from_numpy(io[k]) for k in ("lab", "noise", "neg"))

    def net(x, t, condition=None, **kw):
        return EO.edm_precond_forward(sd, cfg, x, t, condition)

    with torch.no_grad():
        out = net(torch.from_numpy(io["x"]), torch.from_numpy(io["t"]), lab)
        s_cfg = SO.edm_teacher_sample(net, SO.Schedule("edm"), noise, lab, neg, 2.0, 6)
        s_raw = SO.edm_teacher_sample(net, SO.Schedule("edm"), noise, lab, None, None, 6)
    for got, key in ((out, "out"), (s_cfg, "sample_cfg"), (s_raw, "sample")):
        ref = torch.from_numpy(io[key])
        assert ((got - ref).norm() / ref.norm()).item() < 1e-5, key


def test_wan_oracle_matches_reference_glue():
    """oracle/wan_oracle.py against outputs of the UNMODIFIED reference Wan code (classify_forward / prepare / block_forward,
    CausalWanAttnProcessor, inline KV caches, rotary time offset, block-causal mask) run on stand-in diffusers modules
    (oracle/ref_wan_harness.py -> tests/golden/wan_reference_glue.npz).  On the machine that wrote the fixture every
    case but the FlexAttention one is bit-exact; the bounds below only leave room for another host's BLAS summation order
    (fp32: 1e-5; bf16: isolated last-bit flips)."""
    from oracle import wan_oracle as WO

    def close(a, b):
        return torch.equal(a, b) or _rel(a, b) < (2e-3 if a.dtype == torch.bfloat16 else 1e-5)

    g = _load("wan_reference_glue.npz")
    cfg = json.loads(str(g["cfg"]))
    cfg["patch"] = tuple(cfg["patch"])
    sd = WO.make_state_dict(cfg, int(g["seed"]), r_embedder=True)
    x, text, t, r = (torch.from_numpy(g[k]) for k in ("x", "text", "t", "r"))
    B, total, cs = x.shape[0], x.shape[2], 3
    xb = x[:, :, :3]
    variants = {"plain": {}, "r_diff": dict(r=True), "r_abs": dict(r=True, time_cond_type="abs"),
                "r_enc1": dict(r=True, encoder_depth=1), "norm_temb": dict(norm_temb=True),
                "r_enc0_norm": dict(r=True, encoder_depth=0, norm_temb=True), "official": dict(official_sinusoidal=True)}
    with torch.no_grad():
        for name, kw in variants.items():
            kw = dict(kw)
            use_r = kw.pop("r", False)
            out = WO.transformer_forward(sd, cfg, xb, t, text, r_timestep=r if use_r else None, **kw)
            assert close(out, torch.from_numpy(g["bidir_" + name])), name
        sd16 = {k: v.bfloat16() for k, v in sd.items()}
        out = WO.transformer_forward(sd16, cfg, xb.bfloat16(), t, text.bfloat16(), r_timestep=r, norm_temb=True)
        assert close(out, _t(g["bidir_bf16_r_norm"], bf16=True))
        fs = (x.shape[3] // 2) * (x.shape[4] // 2)
        cache = WO.KVCache(cfg, B, total * fs, text.shape[1], torch.float32)
        for ci, (s_, e_) in enumerate(((0, 4), (4, 7))):
            xs = x[:, :, s_:e_]
            ts = torch.full((B, e_ - s_), 937.0, dtype=torch.float64)
            out = WO.transformer_forward(sd, cfg, xs, ts, text, cache=cache, cur_start_frame=s_, store_kv=False)
            assert close(out, torch.from_numpy(g[f"causal_c{ci}_denoise"])), ci
            out = WO.transformer_forward(sd, cfg, xs, torch.zeros(B, e_ - s_, dtype=torch.float64), text, cache=cache,
                                         cur_start_frame=s_, store_kv=True)
            assert close(out, torch.from_numpy(g[f"causal_c{ci}_store"])), ci
        out = WO.transformer_forward(sd, cfg, x, torch.full((B, total), 500.0, dtype=torch.float64), text,
                                     mask=WO.block_causal_mask(total, fs, cs))
        ref = torch.from_numpy(g["causal_block_mask"])
        assert ((out - ref).norm() / ref.norm()).item() < 1e-5
        cache16 = WO.KVCache(cfg, B, 6 * fs, text.shape[1], torch.bfloat16)
        ts = torch.full((B, 3), 937.0, dtype=torch.float64)
        out = WO.transformer_forward(sd16, cfg, x[:, :, :3].bfloat16(), ts, text.bfloat16(), cache=cache16,
                                     cur_start_frame=0, store_kv=True, norm_temb=True)
        assert close(out, _t(g["causal_bf16_norm_store"], bf16=True))
        out = WO.transformer_forward(sd16, cfg, x[:, :, 3:6].bfloat16(), ts, text.bfloat16(), cache=cache16,
                                     cur_start_frame=3, store_kv=False, norm_temb=True)
        assert close(out, _t(g["causal_bf16_norm_c1"], bf16=True))


def test_sdxl_oracle_matches_reference_glue():
    """oracle/sdxl_oracle.py against the UNMODIFIED reference `StableDiffusionXL.forward` + `classify_forward` run on a
    stand-in diffusers UNet (oracle/ref_sdxl_harness.py -> tests/golden/sdxl_reference_glue.npz): conditioning inputs,
    embedding sums, skip routing of the up blocks, output head, eps -> x0 through the reference's noise schedule"""
    from oracle import sdxl_oracle as XO

    g = _load("sdxl_reference_glue.npz")
    cfg = XO.tiny_config()
    sd = XO.make_state_dict(cfg, int(g["seed"]))
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    x, text, pooled, t = (torch.from_numpy(g[k]) for k in ("x", "text", "pooled", "t"))
    with torch.no_grad():
        for pred in ("eps", "x0"):
            out = XO.sdxl_forward(sd, cfg, x, t, text, pooled, fwd_pred_type=pred)
            ref = torch.from_numpy(g[f"f32_{pred}"])
            assert torch.equal(out, ref) or _rel(out, ref) < 1e-5, pred
            out = XO.sdxl_forward(sd16, cfg, x.bfloat16(), t, text.bfloat16(), pooled.bfloat16(), fwd_pred_type=pred,
                                  dtype=torch.bfloat16)
            ref = _t(g[f"bf16_{pred}"], bf16=True)
            assert torch.equal(out, ref) or _rel(out, ref) < 2e-3, pred


def test_sd15_oracle_matches_reference_glue():
    """oracle/sdxl_oracle.py at an SD 1.5-shaped configuration against the UNMODIFIED reference `StableDiffusion15.forward` +
    `classify_forward` with an `encoder_attention_mask` (SD15/network.py:602-678, 46-341) on the stand-in diffusers UNet
    (tests/golden/sd15_reference_glue.npz): the mask -> additive-bias conversion reaching every cross-attention, no addition
    embedding, 1x1-convolution projections, four levels, eps -> x0 through the reference's schedule"""
    from oracle import sdxl_oracle as XO

    g = _load("sd15_reference_glue.npz")
    cfg = XO.tiny_sd15_config()
    sd = XO.make_state_dict(cfg, int(g["seed"]))
    x, text, mask, t = (torch.from_numpy(g[k]) for k in ("x", "text", "mask", "t"))
    with torch.no_grad():
        for pred in ("eps", "x0"):
            out = XO.sd15_forward(sd, cfg, x, t, text, mask, fwd_pred_type=pred)
            ref = torch.from_numpy(g[f"masked_{pred}"])
            assert torch.equal(out, ref) or _rel(out, ref) < 1e-5, pred
        out = XO.sd15_forward(sd, cfg, x, t, text, None, fwd_pred_type="eps")
        assert torch.equal(out, torch.from_numpy(g["nomask_eps"])) or _rel(out, torch.from_numpy(g["nomask_eps"])) < 1e-5
        # the mask matters: the padded samples change, the unpadded one does not
        d = (torch.from_numpy(g["nomask_eps"]) - torch.from_numpy(g["masked_eps"])).flatten(1).abs().amax(1)
        assert d[0] == 0 and d[1] > 1e-2 and d[2] > 1e-2


def test_dit_oracle_matches_reference():
    """oracle/dit_oracle.py against the UNMODIFIED reference DiT (its own __init__, blocks, embedders, modulation, unpatchify,
    SiT convention, r embedder) run with stand-ins for timm's PatchEmbed / Attention / Mlp only
    (oracle/ref_dit_harness.py -> tests/golden/dit_reference.npz)"""
    from oracle import dit_oracle as DO
    from oracle.make_golden import DIT_VARIANTS

    g = _load("dit_reference.npz")
    x, t, r, lab = (torch.from_numpy(g[k]) for k in ("x", "t", "r", "lab"))
    assert torch.equal(DO.pos_embed_2d(128, 4), torch.from_numpy(g["pos_embed"]))
    with torch.no_grad():
        for name, kw in DIT_VARIANTS.items():
            cfg = DO.tiny_config(**kw)
            sd = DO.make_state_dict(cfg, int(g["seed"]))
            rr = r if kw.get("r_timestep") else None
            for pred in ("x0", "flow"):
                out = DO.dit_forward(sd, cfg, x, t, lab, fwd_pred_type=pred, r=rr,
                                     time_cond_type=cfg.get("time_cond_type", "abs"))
                ref = torch.from_numpy(g[f"{name}_{pred}"])
                assert torch.equal(out, ref) or _rel(out, ref) < 1e-5, (name, pred)
