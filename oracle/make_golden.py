"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The reference's own tests hold no golden tensors for this path (shapes
and schedule identities only), so the oracle is pinned against outputs of the reference itself generated here:
weights come from ``oracle.edm_oracle.make_state_dict`` (numpy PCG64, platform independent), inputs from seeded
torch CPU generators and are stored in the fixture, the random draws of the sampler are recorded.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import edm_oracle as EO  # noqa: E402
from oracle.ref_shims import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    "song_tiny": dict(img_resolution=16, img_channels=3, label_dim=10, sigma_shift=0.0, sigma_data=0.5,
                      model_type="SongUNet", augment_dim=9, model_channels=32, channel_mult=[1, 2], num_blocks=1,
                      attn_resolutions=[8], channel_mult_noise=1, embedding_type="positional",
                      encoder_type="standard", decoder_type="standard", resample_filter=[1, 1], dropout=0.0,
                      label_dropout=0, r_timestep=False, drop_precond=None),
    "dhariwal_tiny": dict(img_resolution=16, img_channels=3, label_dim=12, sigma_shift=0.0, sigma_data=0.5,
                          model_type="DhariwalUNet", augment_dim=0, model_channels=64, channel_mult=[1, 2],
                          channel_mult_emb=4, num_blocks=1, attn_resolutions=[8], dropout=0.0, label_dropout=0,
                          r_timestep=False, drop_precond=None),
    "meanflow_tiny": dict(img_resolution=16, img_channels=3, label_dim=10, sigma_shift=0.0, sigma_data=0.5,
                          model_type="SongUNet", augment_dim=0, model_channels=32, channel_mult=[1, 2], num_blocks=1,
                          attn_resolutions=[8], dropout=0.0, label_dropout=0, r_timestep=True, drop_precond="both",
                          net_pred_type="flow", schedule_type="rf"),
}
SEED = 20251018


def _np(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy()
    return t.detach().numpy()


class _Recorder:
    """records torch.randn_like draws and the network outputs during a reference sampler call"""

    def __init__(self, net):
        self.eps, self.outs = [], []
        self._orig = torch.randn_like
        self._net = net

    def __enter__(self):
        def rl(*a, **k):
            v = self._orig(*a, **k)
            self.eps.append(v.clone())
            return v

        torch.randn_like = rl
        self._hook = self._net.register_forward_hook(lambda m, i, o: self.outs.append(o.detach().clone()))
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig
        self._hook.remove()


def build_reference_net(cfg, seed):
    from fastgen.networks.EDM.network import EDMPrecond

    net = EDMPrecond(**cfg)
    sd = EO.make_state_dict(cfg, seed)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("resample_filter" in k for k in missing), missing  # buffers only
    return net.eval(), sd


def edm_fixture(name: str):
    from fastgen.methods.model import FastGenModel

    cfg = CONFIGS[name]
    net, _ = build_reference_net(cfg, SEED)
    g = torch.Generator().manual_seed(SEED + 1)
    B, R = 4, cfg["img_resolution"]
    noise = torch.randn(B, 3, R, R, generator=g)
    labels = torch.eye(cfg["label_dim"])[torch.randint(cfg["label_dim"], (B,), generator=g)]
    t = torch.tensor([80.0, 2.5265068484381206, 0.17256643947993575, 1.1], dtype=torch.float64)
    x_t = noise * t.float().reshape(B, 1, 1, 1)
    fx = dict(cfg=json.dumps(cfg), seed=SEED, noise=_np(noise), labels=_np(labels), t=_np(t), x_t=_np(x_t))
    with torch.no_grad():
        fx["fwd_x0"] = _np(net(x_t, t, condition=labels, fwd_pred_type="x0"))
        fx["fwd_x0_nolabel"] = _np(net(x_t, t, condition=None, fwd_pred_type="x0"))
        fx["fwd_eps"] = _np(net(x_t, t, condition=labels, fwd_pred_type="eps"))
        net64 = net.double()
        fx["fwd_x0_f64"] = _np(net64(x_t.double(), t, condition=labels.double(), fwd_pred_type="x0"))
        net = net.float()
    for steps, stype, tl in [(1, "sde", None), (2, "sde", None), (2, "ode", None), (2, "sde", [80.0, 1.1, 0.0]),
                             (4, "sde", None)]:
        tag = f"gen{steps}_{stype}" + ("_tl" if tl else "")
        torch.manual_seed(SEED + 7)
        with _Recorder(net) as rec:
            out = FastGenModel.generator_fn(net, noise, condition=labels, student_sample_steps=steps,
                                            student_sample_type=stype, t_list=tl)
        fx[tag + "_out"] = _np(out)
        for i, e in enumerate(rec.eps):
            fx[f"{tag}_eps{i}"] = _np(e)
        for i, o in enumerate(rec.outs):
            fx[f"{tag}_x0_{i}"] = _np(o)
    np.savez_compressed(os.path.join(OUT, f"edm_{name}.npz"), **fx)
    print("wrote", name, {k: v.shape for k, v in fx.items() if hasattr(v, "shape") and v.ndim > 1 and "gen" not in k})


def meanflow_fixture():
    from fastgen.methods.consistency_model.mean_flow import MeanFlowModel

    name = "meanflow_tiny"
    cfg = CONFIGS[name]
    net, _ = build_reference_net(cfg, SEED)
    g = torch.Generator().manual_seed(SEED + 2)
    B, R = 4, cfg["img_resolution"]
    noise = torch.randn(B, 3, R, R, generator=g)
    labels = torch.eye(cfg["label_dim"])[torch.randint(cfg["label_dim"], (B,), generator=g)]
    fx = dict(cfg=json.dumps(cfg), seed=SEED, noise=_np(noise), labels=_np(labels))
    t = torch.tensor([0.999, 0.7, 0.3, 0.05], dtype=torch.float64)
    r = torch.tensor([0.5, 0.0, 0.1, 0.0], dtype=torch.float64)
    with torch.no_grad():
        fx["t"], fx["r"] = _np(t), _np(r)
        fx["fwd_flow"] = _np(net(noise, t, condition=labels, r=r, fwd_pred_type="flow"))
        fx["fwd_x0"] = _np(net(noise, t, condition=labels, r=r, fwd_pred_type="x0"))
    for steps, stype in [(1, "sde"), (2, "sde"), (2, "ode"), (4, "ode")]:
        tag = f"gen{steps}_{stype}"
        torch.manual_seed(SEED + 9)
        with _Recorder(net) as rec:
            out = MeanFlowModel.generator_fn(net, noise, condition=labels, student_sample_steps=steps,
                                             student_sample_type=stype)
        fx[tag + "_out"] = _np(out)
        for i, e in enumerate(rec.eps):
            fx[f"{tag}_eps{i}"] = _np(e)
    np.savez_compressed(os.path.join(OUT, f"edm_{name}.npz"), **fx)
    print("wrote", name)


def schedule_fixture():
    from fastgen.networks.noise_schedule import get_noise_schedule

    fx = {}
    edm = get_noise_schedule("edm")
    rf = get_noise_schedule("rf")
    fx["edm_sigmas"] = _np(edm.sigmas)
    for s in (1, 2, 3, 4, 7, 8):
        fx[f"edm_t_list_{s}"] = _np(edm.get_t_list(s))
        fx[f"edm_idx_{s}"] = _np(torch.linspace(edm._max_step, edm._min_step, s + 1).long())
        fx[f"rf_t_list_{s}"] = _np(rf.get_t_list(s))
    fx["rf_rescale"] = _np(rf.rescale_t(rf.get_t_list(4)))
    abar = torch.cumprod(1 - torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2, 0)
    al = get_noise_schedule("alphas", alphas_cumprod=abar)
    fx["alphas_cumprod"] = _np(abar)
    tt = torch.tensor([0.999, 0.74925, 0.4995, 0.24975, 0.0], dtype=torch.float64)
    fx["alphas_t"] = _np(tt)
    fx["alphas_alpha"] = _np(al.alpha(tt))
    fx["alphas_sigma"] = _np(al.sigma(tt))
    fx["alphas_idx"] = _np(al._t_to_idx(tt))
    g = torch.Generator().manual_seed(SEED + 3)
    for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        x = torch.randn(3, 4, 8, 8, generator=g).to(dt)
        e = torch.randn(3, 4, 8, 8, generator=g).to(dt)
        fx[f"x_{dt_name}"], fx[f"e_{dt_name}"] = _np(x), _np(e)
        for sname, sch, t in (("edm", edm, torch.tensor([80.0, 1.1, 0.002], dtype=torch.float64)),
                              ("rf", rf, torch.tensor([0.999, 0.5, 0.0], dtype=torch.float64)),
                              ("alphas", al, torch.tensor([0.999, 0.5, 0.0], dtype=torch.float64))):
            fx[f"{sname}_t3"] = _np(t)
            fx[f"{sname}_latents_{dt_name}"] = _np(sch.latents(x, t[1]))
            fx[f"{sname}_fwd_{dt_name}"] = _np(sch.forward_process(x, e, t))
            fx[f"{sname}_x0_to_eps_{dt_name}"] = _np(sch.x0_to_eps(x, e, t))
            fx[f"{sname}_eps_to_x0_{dt_name}"] = _np(sch.eps_to_x0(x, e, t))
            if sname != "alphas":
                fx[f"{sname}_flow_to_x0_{dt_name}"] = _np(sch.flow_to_x0(x, e, t))
                fx[f"{sname}_x0_to_flow_{dt_name}"] = _np(sch.x0_to_flow(x, e, t))
            else:
                fx[f"{sname}_v_to_x0_{dt_name}"] = _np(sch.v_to_x0(x, e, t))
    # per-frame t ([B,1,T,1,1]) on a video latent (causvid.py:55-56)
    xv = torch.randn(2, 4, 6, 4, 4, generator=g)
    ev = torch.randn(2, 4, 6, 4, 4, generator=g)
    tv = torch.rand(2, 6, generator=g, dtype=torch.float64) * 0.999
    fx["video_x"], fx["video_e"], fx["video_t"] = _np(xv), _np(ev), _np(tv)
    fx["video_fwd"] = _np(rf.forward_process(xv, ev, tv[:, None, :, None, None]))
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **fx)
    print("wrote schedules")


def causal_trace_fixture():
    """call trace of CausVidModel._student_sample_loop against a stub causal network"""
    from fastgen.methods.distribution_matching.causvid import CausVidModel
    from fastgen.networks.network import CausalFastGenNetwork

    trace = []

    class Stub(CausalFastGenNetwork):
        def __init__(self):
            super().__init__(net_pred_type="flow", schedule_type="rf", chunk_size=3, total_num_frames=7)

        def clear_caches(self):
            trace.append(["clear_caches"])

        def forward(self, x_t, t, condition=None, r=None, return_features_early=False, feature_indices=None,
                    return_logvar=False, fwd_pred_type=None, **kw):
            trace.append(["net", list(x_t.shape), float(t[0]), fwd_pred_type, kw.get("cache_tag"),
                          int(kw.get("cur_start_frame")), bool(kw.get("store_kv")), bool(kw.get("is_ar")),
                          sorted(k for k in kw if k not in ("cache_tag", "cur_start_frame", "store_kv", "is_ar"))])
            return 0.5 * x_t

    net = Stub()
    torch.manual_seed(SEED + 11)
    noise = torch.randn(2, 4, 7, 4, 4)
    with _Recorder(net) as rec:
        out = CausVidModel.generator_fn(net, noise, condition=None, neg_condition="neg", student_sample_steps=4,
                                        student_sample_type="sde", t_list=[0.999, 0.937, 0.833, 0.624, 0.0])
    fx = dict(trace=json.dumps(trace), noise=_np(noise), out=_np(out))
    for i, e in enumerate(rec.eps):
        fx[f"eps{i}"] = _np(e)
    np.savez_compressed(os.path.join(OUT, "causal_trace.npz"), **fx)
    print("wrote causal trace:", len(trace), "calls")


def extrapolation_fixture():
    """CausVidModel.generator_fn_extrapolation (causvid.py:187-397) against a stub causal network with a stub VAE:
    call trace, random draws and output of a 3-segment run with 3 overlapping frames and context noise"""
    from fastgen.methods.distribution_matching.causvid import CausVidModel
    from fastgen.networks.network import CausalFastGenNetwork

    trace = []

    class StubVAE:
        def decode(self, z):
            trace.append(["vae.decode", list(z.shape)])
            return 2.0 * z

        def encode(self, v):
            trace.append(["vae.encode", list(v.shape)])
            return 0.5 * v + 0.125

    class Stub(CausalFastGenNetwork):
        def __init__(self):
            super().__init__(net_pred_type="flow", schedule_type="rf", chunk_size=3, total_num_frames=6)
            self.vae = StubVAE()

        def clear_caches(self):
            trace.append(["clear_caches"])

        def forward(self, x_t, t, condition=None, r=None, return_features_early=False, feature_indices=None,
                    return_logvar=False, fwd_pred_type=None, **kw):
            trace.append(["net", list(x_t.shape), float(t[0]), fwd_pred_type, kw.get("cache_tag"),
                          int(kw.get("cur_start_frame")), bool(kw.get("store_kv")), bool(kw.get("is_ar")),
                          sorted(k for k in kw if k not in ("cache_tag", "cur_start_frame", "store_kv", "is_ar"))])
            return 0.5 * x_t + 0.25 * float(t[0])

    net = Stub()
    torch.manual_seed(SEED + 12)
    noise = torch.randn(2, 4, 6, 4, 4)
    with _Recorder(net) as rec:
        out = CausVidModel.generator_fn_extrapolation(net, noise, condition=None, num_segments=3, overlap_frames=3,
                                                      student_sample_steps=2, student_sample_type="sde",
                                                      t_list=[0.999, 0.6, 0.0], context_noise=0.1)
    fx = dict(trace=json.dumps(trace), noise=_np(noise), out=_np(out))
    for i, e in enumerate(rec.eps):
        fx[f"eps{i}"] = _np(e)
    # the ode variant without overlap (fresh noise per segment)
    trace2 = []
    trace, trace_keep = trace2, trace
    torch.manual_seed(SEED + 13)
    with _Recorder(net) as rec2:
        out2 = CausVidModel.generator_fn_extrapolation(net, noise, condition=None, num_segments=2, overlap_frames=0,
                                                       student_sample_steps=2, student_sample_type="ode",
                                                       t_list=[0.999, 0.6, 0.0], context_noise=0)
    fx["out_ode"] = _np(out2)
    for i, e in enumerate(rec2.eps):
        fx[f"ode_eps{i}"] = _np(e)
    np.savez_compressed(os.path.join(OUT, "extrapolation_trace.npz"), **fx)
    print("wrote extrapolation trace:", len(trace_keep), "calls,", len(rec.eps), "draws; ode:", len(rec2.eps), "draws",
          tuple(out.shape), tuple(out2.shape))


def checkpoint_fixture():
    """a `.pth` in the layout of the reference Checkpointer (utils/checkpointer.py:84-96: one torch.save dict) holding the
    state_dict of a real reference EDMPrecond (net + an EMA copy), with one forward of that network for parity"""
    from fastgen.networks.EDM.network import EDMPrecond

    cfg = dict(img_resolution=8, img_channels=3, label_dim=4, model_type="DhariwalUNet", model_channels=32,
               channel_mult=[1], num_blocks=1, attn_resolutions=[8], augment_dim=0, dropout=0.0, label_dropout=0)
    torch.manual_seed(SEED + 21)
    net = EDMPrecond(**cfg).eval()
    with torch.no_grad():
        for prm in net.parameters():       # the reference zero-initialises some layers: make every tensor informative
            if float(prm.abs().max()) < 1e-4:
                prm.normal_(0, 0.05)
    # the EMA entry holds only a few tensors: the reference loads with strict=False (checkpointer.py:157-158)
    ema = {k: v * 0.5 for k, v in list(net.state_dict().items())[:4]}
    save_dict = {"model": {"net": net.state_dict(), "ema": ema}, "optimizer": None, "scheduler": None,
                 "grad_scaler": None, "callbacks": None, "iteration": 7}
    torch.save(save_dict, os.path.join(OUT, "ref_checkpoint_tiny.pth"))
    x = torch.randn(2, 3, 8, 8) * 3
    t = torch.tensor([2.5, 0.3], dtype=torch.float64)
    lab = torch.eye(4)[torch.tensor([1, 3])]
    with torch.no_grad():
        out = net(x, t, condition=lab, fwd_pred_type="x0")
    # teacher sampling with classifier-free guidance (EDM/network.py:976-1026): 6 Euler steps, guidance 2.0 / none
    noise = torch.randn(2, 3, 8, 8)
    neg = torch.zeros(2, 4)
    with torch.no_grad():
        samp_cfg = net.sample(noise, condition=lab, neg_condition=neg, guidance_scale=2.0, num_steps=6)
        samp = net.sample(noise, condition=lab, neg_condition=None, guidance_scale=None, num_steps=6)
    np.savez_compressed(os.path.join(OUT, "ref_checkpoint_tiny_io.npz"), x=_np(x), t=_np(t), lab=_np(lab), out=_np(out),
                        cfg=json.dumps(cfg), noise=_np(noise), neg=_np(neg), sample_cfg=_np(samp_cfg), sample=_np(samp))
    print("wrote checkpoint fixture:", sum(p.numel() for p in net.parameters()), "parameters")


def trig_fixture():
    """TrigNoiseSchedule (noise_schedule.py:1489-1652): scalar functions and the latent conversions, fp32 and bf16"""
    from fastgen.networks.noise_schedule import get_noise_schedule

    ns = get_noise_schedule("trig")
    torch.manual_seed(SEED + 31)
    t = torch.tensor([0.0, 1e-7, 0.3, 0.785, 1.2, torch.pi / 2], dtype=torch.float64)
    fx = dict(t=_np(t), alpha=_np(ns.alpha(t)), sigma=_np(ns.sigma(t)), sqrt_snr=_np(ns.sqrt_snr(t)),
              t_list4=_np(ns.get_t_list(4)), max_sigma=np.float64(ns.max_sigma),
              idx_t=_np(ns.sigma_idx_to_t(torch.tensor([0, 17, 999]))),
              snr_t=_np(ns.sqrt_snr_to_t(torch.tensor([0.1, 1.0, 30.0], dtype=torch.float64))))
    tb = torch.tensor([1.3, 0.4, 1e-9], dtype=torch.float64)
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        xt = torch.randn(3, 4, 8, 8).to(dt)
        v = torch.randn(3, 4, 8, 8).to(dt)
        fx[f"xt_{name}"], fx[f"v_{name}"] = _np(xt), _np(v)
        fx[f"x0_{name}"] = _np(ns.flow_to_x0(xt, v, tb))
        fx[f"flow_{name}"] = _np(ns.x0_to_flow(xt, v, tb))          # (v reused as an x0 input)
        fx[f"fwd_{name}"] = _np(ns.forward_process(xt, v, tb))
    fx["tb"] = _np(tb)
    np.savez_compressed(os.path.join(OUT, "schedule_trig.npz"), **fx)
    print("wrote trig schedule fixture")


def wan_glue_fixture():
    """Outputs of the UNMODIFIED reference Wan glue code (fastgen/networks/Wan/network.py + network_causal.py) run on
    stand-in diffusers modules (oracle/ref_wan_harness.py): bidirectional forward with every r-embedder / norm_temb /
    sinusoid variant, causal KV-cached chunks, the block-causal mask (eager FlexAttention)."""
    os.environ["FASTGEN_DISABLE_FLEX_ATTENTION"] = "0"
    os.environ["FASTGEN_FLEX_COMPILE"] = "0"
    from oracle import ref_wan_harness as H
    from oracle import wan_oracle as WO

    cfg = WO.tiny_config(layers=2)
    sd = WO.make_state_dict(cfg, 11, r_embedder=True)
    g = torch.Generator().manual_seed(SEED + 41)
    B, total, cs, Hh, Ww = 2, 7, 3, 8, 8
    x = torch.randn(B, 16, total, Hh, Ww, generator=g)
    text = torch.randn(B, 24, cfg["text_dim"], generator=g)
    t = torch.tensor([800.0, 500.0], dtype=torch.float64)
    r = torch.tensor([300.0, 500.0], dtype=torch.float64)
    fx = dict(cfg=json.dumps(cfg), seed=np.int64(11), x=_np(x), text=_np(text), t=_np(t), r=_np(r))
    variants = {"plain": {}, "r_diff": dict(r=True), "r_abs": dict(r=True, time_cond_type="abs"),
                "r_enc1": dict(r=True, encoder_depth=1), "norm_temb": dict(norm_temb=True),
                "r_enc0_norm": dict(r=True, encoder_depth=0, norm_temb=True), "official": dict(official_sinusoidal=True)}
    xb = x[:, :, :3]
    with torch.no_grad():
        for name, kw in variants.items():
            kw = dict(kw)
            use_r = kw.pop("r", False)
            tr = H.build_bidirectional(cfg, sd, r_timestep=True, **kw)
            fx["bidir_" + name] = _np(tr(hidden_states=xb, timestep=t, encoder_hidden_states=text,
                                         r_timestep=r if use_r else None, feature_indices=set()))
        sd16 = {k: v.bfloat16() for k, v in sd.items()}
        tr = H.build_bidirectional(cfg, sd16, r_timestep=True, norm_temb=True).to(torch.bfloat16)
        fx["bidir_bf16_r_norm"] = _np(tr(hidden_states=xb.bfloat16(), timestep=t, encoder_hidden_states=text.bfloat16(),
                                         r_timestep=r, feature_indices=set()))
        # causal: chunk 0 = frames 0..3 (remainder absorbed), chunk 1 = frames 4..6; denoise pass then cache-fill pass
        trc = H.build_causal(cfg, sd)
        for ci, (s_, e_) in enumerate(((0, 4), (4, 7))):
            xs = x[:, :, s_:e_]
            ts = torch.full((B, e_ - s_), 937.0, dtype=torch.float64)
            fx[f"causal_c{ci}_denoise"] = _np(H.causal_call(trc, xs, ts, text, chunk_size=cs, total_num_frames=total,
                                                            cur_start_frame=s_, store_kv=False, is_ar=True))
            fx[f"causal_c{ci}_store"] = _np(H.causal_call(trc, xs, torch.zeros(B, e_ - s_, dtype=torch.float64), text,
                                                          chunk_size=cs, total_num_frames=total, cur_start_frame=s_,
                                                          store_kv=True, is_ar=True))
        trm = H.build_causal(cfg, sd)
        fx["causal_block_mask"] = _np(H.causal_call(trm, x, torch.full((B, total), 500.0, dtype=torch.float64), text,
                                                    chunk_size=cs, total_num_frames=total, is_ar=False))
        trn = H.build_causal(cfg, sd16, norm_temb=True).to(torch.bfloat16)
        ts = torch.full((B, 3), 937.0, dtype=torch.float64)
        fx["causal_bf16_norm_store"] = _np(H.causal_call(trn, x[:, :, :3].bfloat16(), ts, text.bfloat16(), chunk_size=cs,
                                                         total_num_frames=6, cur_start_frame=0, store_kv=True, is_ar=True))
        fx["causal_bf16_norm_c1"] = _np(H.causal_call(trn, x[:, :, 3:6].bfloat16(), ts, text.bfloat16(), chunk_size=cs,
                                                      total_num_frames=6, cur_start_frame=3, store_kv=False, is_ar=True))
    np.savez_compressed(os.path.join(OUT, "wan_reference_glue.npz"), **fx)
    print("wrote Wan reference-glue fixture:", len(fx), "arrays")


def sdxl_glue_fixture():
    """Outputs of the UNMODIFIED reference `StableDiffusionXL.forward` + `classify_forward` (SDXL/network.py:334-416,
    SD15/network.py:46-341) on a stand-in diffusers UNet (oracle/ref_sdxl_harness.py), fp32 and bf16, eps and x0"""
    from oracle import ref_sdxl_harness as H
    from oracle import sdxl_oracle as XO

    cfg = XO.tiny_config()
    sd = XO.make_state_dict(cfg, 21)
    g = torch.Generator().manual_seed(SEED + 51)
    x = torch.randn(2, 4, 32, 32, generator=g)
    text = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g)
    pooled = torch.randn(2, cfg["pooled_dim"], generator=g)
    t = torch.tensor([0.999, 0.4], dtype=torch.float64)
    fx = dict(seed=np.int64(21), x=_np(x), text=_np(text), pooled=_np(pooled), t=_np(t))
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    with torch.no_grad():
        for pred in ("eps", "x0"):
            fx[f"f32_{pred}"] = _np(H.reference_sdxl_forward(cfg, sd, x, t, text, pooled, pred))
            fx[f"bf16_{pred}"] = _np(H.reference_sdxl_forward(cfg, sd16, x.bfloat16(), t, text.bfloat16(), pooled.bfloat16(), pred))
    np.savez_compressed(os.path.join(OUT, "sdxl_reference_glue.npz"), **fx)
    print("wrote SDXL reference-glue fixture")


def sd15_glue_fixture():
    """Outputs of the UNMODIFIED reference `StableDiffusion15.forward` + `classify_forward` with an `encoder_attention_mask`
    (SD15/network.py:602-678, 46-341: the mask becomes the additive bias (1 - mask) * -10000 handed to every block) on a
    stand-in diffusers UNet at an SD 1.5-shaped configuration: no addition embedding, 1x1-conv proj_in / proj_out, four
    levels, heads of 8 / 16 / 32 channels"""
    from oracle import ref_sdxl_harness as H
    from oracle import sdxl_oracle as XO

    cfg = XO.tiny_sd15_config()
    sd = XO.make_state_dict(cfg, 23)
    g = torch.Generator().manual_seed(SEED + 53)
    x = torch.randn(3, 4, 16, 16, generator=g)
    text = torch.randn(3, 77, cfg["cross_attention_dim"], generator=g)
    mask = torch.zeros(3, 77)
    for b, n in enumerate((77, 9, 31)):
        mask[b, :n] = 1
    t = torch.tensor([0.999, 0.4, 0.05], dtype=torch.float64)
    fx = dict(seed=np.int64(23), x=_np(x), text=_np(text), mask=_np(mask), t=_np(t))
    with torch.no_grad():
        for pred in ("eps", "x0"):
            fx[f"masked_{pred}"] = _np(H.reference_sd15_forward(cfg, sd, x, t, text, mask, pred))
        fx["nomask_eps"] = _np(H.reference_sd15_forward(cfg, sd, x, t, text, None, "eps"))
    np.savez_compressed(os.path.join(OUT, "sd15_reference_glue.npz"), **fx)
    print("wrote SD 1.5 reference-glue fixture")


DIT_VARIANTS = {"d64": {}, "d72": dict(hidden_size=144, num_heads=2, depth=3), "sit": dict(use_sit_convention=True),
                "r_diff": dict(r_timestep=True, time_cond_type="diff"), "r_abs": dict(r_timestep=True, time_cond_type="abs")}


def dit_fixture():
    """Outputs of the UNMODIFIED reference DiT (fastgen/networks/DiT/network.py) with stand-ins for timm's PatchEmbed /
    Attention / Mlp (oracle/ref_dit_harness.py): plain, 72-wide heads, SiT convention, r embedder (diff / abs)"""
    from oracle import dit_oracle as DO
    from oracle import ref_dit_harness as H

    g = torch.Generator().manual_seed(SEED + 61)
    cfg0 = DO.tiny_config()
    x = torch.randn(3, cfg0["in_channels"], cfg0["input_size"], cfg0["input_size"], generator=g)
    t = torch.tensor([0.9, 0.5, 0.1], dtype=torch.float64)
    r = torch.tensor([0.2, 0.5, 0.05], dtype=torch.float64)
    lab = torch.eye(cfg0["num_classes"])[torch.tensor([1, 7, 3])]
    lab[2] = 0          # all-zero row -> the null class
    fx = dict(x=_np(x), t=_np(t), r=_np(r), lab=_np(lab), seed=np.int64(31))
    with torch.no_grad():
        for name, kw in DIT_VARIANTS.items():
            cfg = DO.tiny_config(**kw)
            net = H.build_reference_dit(cfg, DO.make_state_dict(cfg, 31))
            rr = r if kw.get("r_timestep") else None
            for pred in ("x0", "flow"):
                fx[f"{name}_{pred}"] = _np(net(x, t, condition=lab, r=rr, fwd_pred_type=pred))
            if name == "d64":
                fx["pos_embed"] = _np(net.pos_embed)
    np.savez_compressed(os.path.join(OUT, "dit_reference.npz"), **fx)
    print("wrote DiT reference fixture:", len(fx), "arrays")


def facebook_dit_keys(sd):
    """a Facebook-DiT-named view of a state dict in the reference's naming (the inverse of DiT/network.py:375-385)"""
    fb = {}
    for k, v in sd.items():
        k = k.replace("t_embedder.proj_net", "t_embedder.mlp").replace(".attention.", ".attn.")
        k = k.replace(".feed_forward.", ".mlp.").replace("y_embedder.class_embeddings", "y_embedder.embedding_table")
        k = k.replace("final_layer.projection", "final_layer.linear")
        k = k.replace("final_layer.adaptive_params", "final_layer.adaLN_modulation")
        k = k.replace(".conditioning_net.", ".adaLN_modulation.")
        fb[k] = v
    return fb


def dit_remap_fixture():
    """The UNMODIFIED reference `DiT.load_state_dict` (DiT/network.py:346-401) on a Facebook-DiT-named checkpoint whose
    tensors are tagged with their own index: which tensor ends up under which reference key, the returned missing /
    unexpected lists, and the pos_embed buffer handling (matching and mismatching shape)."""
    from oracle import dit_oracle as DO
    from oracle import ref_dit_harness as H

    cfg = DO.tiny_config()
    net = H.build_reference_dit(cfg, DO.make_state_dict(cfg, 31))
    ref_sd = net.state_dict()
    fb = facebook_dit_keys(ref_sd)
    tagged = {k: torch.full_like(v, float(i + 1)) for i, (k, v) in enumerate(fb.items())}
    res = net.load_state_dict(tagged, strict=False)
    after = net.state_dict()
    fx = dict(src=list(tagged.keys()), missing=list(res.missing_keys), unexpected=list(res.unexpected_keys),
              tag_of={k: float(v.flatten()[0]) for k, v in after.items()})
    bad = dict(tagged)
    bad["pos_embed"] = torch.full((1, 3, 5), 99.0)
    res2 = net.load_state_dict(bad, strict=False)
    fx["bad_pos_missing"], fx["bad_pos_tag"] = list(res2.missing_keys), float(net.pos_embed.flatten()[0])
    with open(os.path.join(OUT, "dit_facebook_key_remap.json"), "w") as f:
        json.dump(fx, f, indent=0)
    print("dit_facebook_key_remap.json", len(fx["src"]), "keys, missing", fx["missing"], fx["bad_pos_missing"], fx["bad_pos_tag"])


def wan_sample_stub_flow(x, t, condition, cur_start_frame=0, store_kv=False, cache_state=0.0):
    """deterministic stand-in for a Wan forward (flow prediction): depends on the latent, the timestep, the condition, the
    chunk position and on what the cache held -- so a wrong call order, cache tag or argument shows in the samples.
    Shared by the fixture generator, the oracle restatement and the CPU test of the product loop."""
    tt = t.float().reshape(t.shape[0], *([1] * (x.dim() - 1)))
    c = condition.float().mean(dim=tuple(range(1, condition.dim()))).reshape(-1, *([1] * (x.dim() - 1)))
    out = 0.6 * x.float() + 0.25 * torch.sin(3.0 * tt) + 0.5 * c + 0.01 * cur_start_frame + 0.05 * cache_state
    return out.to(x.dtype)


def wan_sample_fixture():
    """Traces and samples of the UNMODIFIED reference teacher samplers `Wan.sample` (Wan/network.py:919-988) and
    `CausalWan.sample` (network_causal.py:1186-1295), called as plain functions on a stand-in `self`: a stub network
    (wan_sample_stub_flow), the reference's own rf noise schedule, and fastgen_b200.unipc.FlowUniPCMultistepScheduler
    standing in for the diffusers scheduler object (the loop only uses its interface; the scheduler arithmetic is not
    what this fixture pins)."""
    from fastgen.networks.noise_schedule import get_noise_schedule
    from fastgen_b200.unipc import FlowUniPCMultistepScheduler
    from oracle import ref_wan_harness as H

    RW, RC = H.load_reference_wan()

    class Stub:
        schedule_type = "rf"
        training = False

        def __init__(self, chunk_size=3):
            self.noise_scheduler = get_noise_schedule("rf")
            self.unipc_scheduler = FlowUniPCMultistepScheduler()
            self.chunk_size = chunk_size
            self.trace, self.cache = [], {}

        def eval(self):
            self.trace.append(("eval",))

        def train(self, mode=True):
            self.trace.append(("train", bool(mode)))

        def clear_caches(self):
            self.trace.append(("clear_caches",))
            self.cache = {}

        def __call__(self, x, t, condition=None, cache_tag=None, cur_start_frame=0, store_kv=False, is_ar=None, **kw):
            extra = {k: v for k, v in kw.items() if v not in (None, False) and v != {}}
            self.trace.append(("fwd", [float(v) for v in t.float().flatten()], str(t.dtype), float(condition.flatten()[0]),
                               cache_tag, int(cur_start_frame), bool(store_kv), is_ar, sorted(extra)))
            state = self.cache.get(cache_tag, 0.0)
            if store_kv:
                self.cache[cache_tag] = state + float(x.float().mean())
            return wan_sample_stub_flow(x, t, condition, cur_start_frame, store_kv, state)

    g = torch.Generator().manual_seed(SEED + 77)
    fx = {}
    cond, neg = torch.randn(2, 5, 8, generator=g), torch.randn(2, 5, 8, generator=g)
    fx["cond"], fx["neg"] = _np(cond), _np(neg)
    with torch.no_grad():
        for tag, dtype, kw in (("bidir_cfg", torch.float32, dict(guidance_scale=5.0, num_steps=6, shift=5.0)),
                               ("bidir_nocfg", torch.float32, dict(guidance_scale=None, num_steps=4, shift=3.0)),
                               ("bidir_bf16", torch.bfloat16, dict(guidance_scale=3.0, num_steps=5, shift=5.0))):
            noise = torch.randn(2, 4, 3, 4, 4, generator=g).to(dtype)
            st = Stub()
            out = RW.Wan.sample(st, noise.clone(), condition=cond.to(dtype), neg_condition=neg.to(dtype), **kw)
            fx[tag + "_noise"], fx[tag + "_out"], fx[tag + "_trace"] = _np(noise), _np(out), json.dumps(st.trace)
            fx[tag + "_kw"] = json.dumps(kw)
        for tag, frames, kw in (("causal_cfg", 7, dict(guidance_scale=4.0, sample_steps=4, shift=5.0)),
                                ("causal_ctx", 6, dict(guidance_scale=None, sample_steps=3, shift=3.0, context_noise=0.1)),
                                ("causal_short", 2, dict(guidance_scale=2.0, sample_steps=3, shift=5.0))):
            noise = torch.randn(2, 4, frames, 4, 4, generator=g)
            draws = []
            orig = torch.randn_like

            def rl(x, **k2):
                e = torch.randn(x.shape, generator=g).to(x.dtype)
                draws.append(e)
                return e

            torch.randn_like = rl
            try:
                st = Stub()
                out = RC.CausalWan.sample(st, noise.clone(), condition=cond, neg_condition=neg, **kw)
            finally:
                torch.randn_like = orig
            fx[tag + "_noise"], fx[tag + "_out"], fx[tag + "_trace"] = _np(noise), _np(out), json.dumps(st.trace)
            fx[tag + "_kw"] = json.dumps(kw)
            for i, e in enumerate(draws):
                fx[f"{tag}_eps{i}"] = _np(e)
    np.savez_compressed(os.path.join(OUT, "wan_sample_trace.npz"), **fx)
    print("wan_sample_trace.npz", {k: v.shape for k, v in fx.items() if hasattr(v, "shape") and v.ndim > 0 and "out" in k})


def original_wan_keys(layers: int = 2):
    """key names of an original-format (Wan-AI/Wan2.1 `*.pth`-style) checkpoint of a T2V model"""
    keys = ["patch_embedding.weight", "patch_embedding.bias", "head.head.weight", "head.head.bias", "head.modulation"]
    for i in (0, 2):
        keys += [f"text_embedding.{i}.weight", f"text_embedding.{i}.bias", f"time_embedding.{i}.weight",
                 f"time_embedding.{i}.bias"]
    keys += ["time_projection.1.weight", "time_projection.1.bias"]
    for b in range(layers):
        p = f"blocks.{b}."
        keys += [p + "modulation", p + "norm3.weight", p + "norm3.bias", p + "ffn.0.weight", p + "ffn.0.bias",
                 p + "ffn.2.weight", p + "ffn.2.bias"]
        for a in ("self_attn", "cross_attn"):
            for part in ("q", "k", "v", "o"):
                keys += [f"{p}{a}.{part}.weight", f"{p}{a}.{part}.bias"]
            keys += [f"{p}{a}.norm_q.weight", f"{p}{a}.norm_k.weight"]
    return keys


def wan_remap_fixture():
    """What the UNMODIFIED reference `Wan.load_state_dict` (Wan/network.py:990-1075) hands to nn.Module.load_state_dict for
    original-format checkpoints: flat or nested under 'generator' / 'state_dict' / 'model', keys with and without the
    'model.' / 'module.' prefixes.  The instance is created without __init__ (no diffusers model is needed for the key
    logic) and nn.Module.load_state_dict is replaced by a recorder for the call."""
    from oracle import ref_wan_harness as H

    RW, _ = H.load_reference_wan()
    obj = RW.Wan.__new__(RW.Wan)
    torch.nn.Module.__init__(obj)
    obj._use_wan_official_sinusoidal = True
    got = {}
    orig = torch.nn.Module.load_state_dict
    torch.nn.Module.load_state_dict = lambda self, sd, **kw: got.update(keys=list(sd.keys()), kw=sorted(kw))
    fx = {}
    try:
        for tag, nest, prefix in (("flat", None, ""), ("generator_model", "generator", "model."),
                                  ("state_dict_module", "state_dict", "module."), ("model_plain", "model", "")):
            src = {prefix + k: torch.zeros(1) for k in original_wan_keys()}
            RW.Wan.load_state_dict(obj, {nest: src} if nest else src, strict=False)
            fx[tag] = dict(nest=nest, src=list(src.keys()), dst=got["keys"], kw=got["kw"])
        # a checkpoint that already carries diffusers names passes through untouched
        src = {"transformer.blocks.0.attn1.to_q.weight": torch.zeros(1), "transformer.scale_shift_table": torch.zeros(1)}
        RW.Wan.load_state_dict(obj, src)
        fx["diffusers_names"] = dict(nest=None, src=list(src.keys()), dst=got["keys"], kw=got["kw"])
    finally:
        torch.nn.Module.load_state_dict = orig
    with open(os.path.join(OUT, "wan_original_key_remap.json"), "w") as f:
        json.dump(fx, f, indent=0)
    print("wan_original_key_remap.json", {k: len(v["dst"]) for k, v in fx.items()})


def handoff_fixture():
    """tests/golden/handoff_reference.npz: the UNMODIFIED reference wrappers WanTextEncoder / WanVideoEncoder
    (fastgen/networks/Wan/network.py:431-539) and SDXLTextEncoder / SDXLImageEncoder (SDXL/network.py:46-156), built with
    __new__ (their constructors download checkpoints) around the stand-in networks of oracle/ref_handoff_harness.py."""
    from oracle import ref_handoff_harness as H
    from oracle.ref_sdxl_harness import load_reference_sdxl
    from oracle.ref_wan_harness import load_reference_wan

    wan, _ = load_reference_wan()
    sdxl = load_reference_sdxl()
    out = {}
    g = torch.Generator().manual_seed(SEED + 77)
    # --- Wan text encoder: real transformers UMT5 (toy size), prompt cleaning, truncation to the token count, zero padding
    te = wan.WanTextEncoder.__new__(wan.WanTextEncoder)
    te.text_encoder, te.tokenizer = H.toy_umt5(), H.ToyTokenizer(97)
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        out[f"wan_text_{name}"] = _np(te.encode(H.PROMPTS, precision=dt))
    out["wan_text_single_f32"] = _np(te.encode(H.PROMPTS[1]))
    # --- Wan VAE wrapper: latent (de)normalisation, dtype hand-over, clip
    ve = wan.WanVideoEncoder.__new__(wan.WanVideoEncoder)
    ve.vae = H.ToyWanVAE().eval().requires_grad_(False)
    lat32 = torch.randn(2, 16, 3, 6, 10, generator=g) * 1.3
    lat16 = lat32.bfloat16()
    video = torch.randn(1, 3, 5, 16, 24, generator=g)
    out["wan_latent_f32"], out["wan_latent_bf16"], out["wan_video"] = _np(lat32), _np(lat16), _np(video)
    out["wan_decode_f32"] = _np(ve.decode(lat32))
    out["wan_decode_bf16"] = _np(ve.decode(lat16))          # bf16 latent, fp32 VAE: coefficients and ops round in bf16
    out["wan_encode_sample"] = _np(ve.encode(video, mode="sample"))
    out["wan_encode_argmax"] = _np(ve.encode(video.bfloat16(), mode="argmax"))
    # --- SDXL text encoders: real transformers CLIP pair (toy size)
    st = sdxl.SDXLTextEncoder.__new__(sdxl.SDXLTextEncoder)
    st.text_encoder_one, st.text_encoder_two = H.toy_clip_pair()
    st.tokenizer_one, st.tokenizer_two = H.ToyTokenizer(97), H.ToyTokenizer(97)
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        pe, pooled = st.encode(H.PROMPTS, precision=dt)
        out[f"sdxl_text_{name}"], out[f"sdxl_pooled_{name}"] = _np(pe), _np(pooled)
    # --- SDXL VAE wrapper
    ie = sdxl.SDXLImageEncoder.__new__(sdxl.SDXLImageEncoder)
    ie.vae = H.ToySDVAE().eval().requires_grad_(False)
    slat = torch.randn(2, 4, 8, 8, generator=g) * 2.0
    img = torch.randn(2, 3, 32, 32, generator=g)
    out["sdxl_latent"], out["sdxl_image"] = _np(slat), _np(img)
    out["sdxl_decode_f32"] = _np(ie.decode(slat))
    out["sdxl_encode_f32"] = _np(ie.encode(img))
    ie.vae = H.ToySDVAE().eval().requires_grad_(False).bfloat16()
    out["sdxl_decode_bf16"] = _np(ie.decode(slat.bfloat16()))
    np.savez_compressed(os.path.join(OUT, "handoff_reference.npz"), **out)
    print("handoff_reference.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    schedule_fixture()
    edm_fixture("song_tiny")
    edm_fixture("dhariwal_tiny")
    meanflow_fixture()
    causal_trace_fixture()
    extrapolation_fixture()
    checkpoint_fixture()
    trig_fixture()
    wan_glue_fixture()
    sdxl_glue_fixture()
    sd15_glue_fixture()
    dit_fixture()
    wan_sample_fixture()
    wan_remap_fixture()
    dit_remap_fixture()
    handoff_fixture()
