"""Parity of the CUDA Wan path (through the C ABI) against the CPU oracle (oracle/wan_oracle.py: pinned bit-exactly to the
reference's own Wan code run on stand-in diffusers modules, tests/golden/wan_reference_glue.npz; the diffusers module
internals are restated, see the oracle header) on identical random-init weights, seeds and inputs.

Tolerance (BASELINE.json north star): per-step x0 relative L2 <= 1e-2 on the bf16 path.  The oracle is evaluated both
in bfloat16 (what the reference network computes) and in float32 on the same bf16-rounded weights (the exact value)."""
import pytest
import torch

from oracle import sampler_oracle as SO
from oracle import wan_oracle as WO

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _nets(cfg, seed, cls_name, **kw):
    import fastgen_b200.wan as W

    sd = {k: v.bfloat16() for k, v in WO.make_state_dict(cfg, seed, r_embedder=bool(kw.get("r_timestep"))).items()}
    net = getattr(W, cls_name)(transformer_config=cfg, enable_logvar_linear=False, **kw)
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return net.to(torch.bfloat16).cuda().eval(), sd


def _inputs(cfg, b, f, h, w, text_len=24, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, cfg["in_channels"], f, h, w, generator=g).bfloat16()
    text = torch.randn(b, text_len, cfg["text_dim"], generator=g).bfloat16()
    return x, text


@pytest.mark.parametrize("official", [False, True])
def test_bidirectional_forward_matches_oracle(official):
    cfg = WO.tiny_config()
    net, sd = _nets(cfg, 0, "Wan", use_wan_official_sinusoidal=official)
    x, text = _inputs(cfg, 2, 3, 8, 12)
    t = torch.tensor([0.8, 0.3], dtype=torch.float64)
    for pred in ("x0", "flow"):
        out = net(x.cuda(), t.cuda(), condition=text.cuda(), fwd_pred_type=pred)
        assert out.shape == x.shape and out.dtype == x.dtype
        ref16 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type=pred, dtype=torch.bfloat16, official_sinusoidal=official)
        ref32 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type=pred, dtype=torch.float32, official_sinusoidal=official)
        assert _rel(out, ref32) < TOL, (pred, _rel(out, ref32))
        assert _rel(out, ref16) < 2 * TOL, (pred, _rel(out, ref16))
        # and the CUDA path is at least as close to the exact value as the bf16 reference arithmetic is
        assert _rel(out, ref32) < 1.5 * _rel(ref16, ref32) + 1e-3


@pytest.mark.parametrize("kw", [dict(), dict(time_cond_type="abs"), dict(encoder_depth=1), dict(norm_temb=True),
                                dict(encoder_depth=0, norm_temb=True)],
                         ids=["diff", "abs", "encoder_depth1", "norm_temb", "encoder_depth0_norm_temb"])
def test_r_embedder_variants_match_oracle(kw):
    """second timestep r (MeanFlow on Wan): r embedder added to / replacing the time projection, sCM-style normalised
    AdaLN vectors -- reference classify_forward_prepare (Wan/network.py:330-353) and block_forward (:131-136)"""
    cfg = WO.tiny_config(layers=3)
    net, sd = _nets(cfg, 4, "Wan", r_timestep=True, **kw)
    x, text = _inputs(cfg, 2, 3, 8, 12)
    t = torch.tensor([0.8, 0.5], dtype=torch.float64)
    r = torch.tensor([0.3, 0.5], dtype=torch.float64)
    out = net(x.cuda(), t.cuda(), condition=text.cuda(), r=r.cuda(), fwd_pred_type="flow")
    ref32 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type="flow", dtype=torch.float32, r=r, **kw)
    assert _rel(out, ref32) < TOL, _rel(out, ref32)
    # without r the r embedder is not used at all (norm_temb still is)
    out0 = net(x.cuda(), t.cuda(), condition=text.cuda(), fwd_pred_type="flow")
    ref0 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type="flow", dtype=torch.float32,
                          norm_temb=kw.get("norm_temb", False))
    assert _rel(out0, ref0) < TOL, _rel(out0, ref0)
    assert _rel(out0, ref32) > 5 * TOL      # ... and r does change the prediction


def test_causal_r_embedder_per_frame():
    """CausalWan with r (network_causal.py:641-665): per-frame r embedding, normalised per-frame AdaLN vectors"""
    cfg = WO.tiny_config(layers=2)
    total = 4
    net, sd = _nets(cfg, 5, "CausalWan", total_num_frames=total, chunk_size=2, r_timestep=True, norm_temb=True)
    x, text = _inputs(cfg, 2, total, 8, 8)
    t = torch.tensor([0.7, 0.4], dtype=torch.float64)
    r = torch.tensor([0.2, 0.1], dtype=torch.float64)
    out = net(x.cuda(), t.cuda(), condition=text.cuda(), r=r.cuda(), fwd_pred_type="flow", is_ar=True)
    B = x.shape[0]
    ts = (t * 1000.0).view(-1, 1).expand(B, total)
    rs = (r * 1000.0).view(-1, 1).expand(B, total)
    ref = WO.transformer_forward(sd, cfg, x.float(), ts, text.float(), r_timestep=rs, norm_temb=True)
    assert _rel(out, ref) < TOL, _rel(out, ref)


def test_causal_chunks_with_kv_cache_match_oracle():
    """chunk 0 (frames 0-3, remainder absorbed) denoise + store, chunk 1 (frames 4-6) attends to the cached prefix"""
    cfg = WO.tiny_config(layers=3)
    total = 7
    net, sd = _nets(cfg, 1, "CausalWan", total_num_frames=total, chunk_size=3)
    x, text = _inputs(cfg, 2, total, 8, 8)
    sd32 = sd
    cache16 = WO.KVCache(cfg, 2, total * 16, text.shape[1], torch.bfloat16)
    cache32 = WO.KVCache(cfg, 2, total * 16, text.shape[1], torch.float32)
    net.clear_caches()
    for (s, e) in [(0, 4), (4, 7)]:
        xc = x[:, :, s:e].contiguous()
        for tval, store in [(0.937, False), (0.0, True)]:
            t = torch.full((2,), tval, dtype=torch.float64)
            out = net(xc.cuda(), t.cuda(), condition=text.cuda(), fwd_pred_type="x0", cache_tag="pos", cur_start_frame=s,
                      store_kv=store, is_ar=True, neg_condition=None)
            r16 = WO.causal_wan_forward(sd, cfg, xc, t, text, cache=cache16, cur_start_frame=s, store_kv=store,
                                        total_num_frames=total, is_ar=True, dtype=torch.bfloat16)
            r32 = WO.causal_wan_forward(sd32, cfg, xc, t, text, cache=cache32, cur_start_frame=s, store_kv=store,
                                        total_num_frames=total, is_ar=True, dtype=torch.float32)
            assert _rel(out, r32) < TOL, (s, tval, _rel(out, r32))
            assert _rel(out, r16) < 2 * TOL, (s, tval, _rel(out, r16))
    # the cache contents themselves (layer 0, roped keys) match the oracle's
    kv = net._kv["pos"]
    k_ref = cache32.k[0].flatten(2)
    assert _rel(kv["k"][0], k_ref) < 1e-2


def test_causal_full_sequence_uses_block_causal_mask():
    cfg = WO.tiny_config()
    total = 5
    net, sd = _nets(cfg, 2, "CausalWan", total_num_frames=total, chunk_size=2)
    x, text = _inputs(cfg, 1, total, 8, 8)
    t = torch.tensor([[0.9, 0.9, 0.9, 0.5, 0.5]], dtype=torch.float64)   # per-frame t (diffusion forcing)
    out = net(x.cuda(), t.cuda(), condition=text.cuda(), fwd_pred_type="x0")
    r32 = WO.causal_wan_forward(sd, cfg, x, t, text, total_num_frames=total, chunk_size=2, dtype=torch.float32)
    assert _rel(out, r32) < TOL, _rel(out, r32)
    # is_ar=True on the full sequence: no mask (reference: current_block_mask = None if is_ar)
    out_ar = net(x.cuda(), t.cuda(), condition=text.cuda(), fwd_pred_type="x0", is_ar=True)
    r32_ar = WO.causal_wan_forward(sd, cfg, x, t, text, total_num_frames=total, chunk_size=2, is_ar=True, dtype=torch.float32)
    assert _rel(out_ar, r32_ar) < TOL
    assert _rel(out, r32_ar) > 1e-2  # and the mask matters


@pytest.mark.parametrize("stype", ["sde", "ode"])
def test_self_forcing_sampler_matches_oracle(stype):
    """CausVidModel.generator_fn (4-step, chunked, KV cache) end to end against the oracle loop"""
    from fastgen_b200.sampler import CausVidModel

    cfg = WO.tiny_config()
    total, chunk = 7, 3
    net, sd = _nets(cfg, 3, "CausalWan", total_num_frames=total, chunk_size=chunk, use_wan_official_sinusoidal=True)
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(1, 16, total, 8, 8, generator=g).bfloat16()
    text = torch.randn(1, 16, cfg["text_dim"], generator=g).bfloat16()
    tl = [0.999, 0.937, 0.833, 0.624, 0.0]
    chunks = SO.causal_chunks(total, chunk)
    eps = [torch.randn(1, 16, e - s, 8, 8, generator=g).bfloat16() for (s, e) in chunks for _ in range(3)]
    sched = SO.Schedule("rf")
    cache = WO.KVCache(cfg, 1, total * 16, text.shape[1], torch.float32)

    def onet(xc, tb, cur_start_frame, store_kv):
        return WO.causal_wan_forward(sd, cfg, xc, tb, text, cache=cache, cur_start_frame=cur_start_frame,
                                     store_kv=store_kv, total_num_frames=total, is_ar=True, dtype=torch.float32,
                                     official_sinusoidal=True)

    lat = sched.latents(noise, torch.tensor(tl[0], dtype=torch.float64))
    ref = SO.causvid_sample_loop(onet, sched, lat, torch.tensor(tl, dtype=torch.float64), chunk,
                                 eps if stype == "sde" else [], stype)
    it = iter(eps)
    orig = torch.randn_like
    torch.randn_like = lambda x, **kw: next(it).to(x.device, x.dtype)
    try:
        out = CausVidModel.generator_fn(net, noise.cuda(), student_sample_steps=4, t_list=tl, condition=text.cuda(),
                                        student_sample_type=stype, neg_condition=None)
    finally:
        torch.randn_like = orig
    assert out.shape == noise.shape and out.dtype == noise.dtype
    assert _rel(out, ref) < 2e-2, _rel(out, ref)   # 4 re-noised steps x 2 chunks of bf16 latents


def test_state_dict_keys_and_original_wan_remap():
    import fastgen_b200.wan as W

    cfg = WO.tiny_config(layers=1)
    net = W.Wan(transformer_config=cfg, use_wan_official_sinusoidal=True)
    keys = set(net.state_dict().keys())
    assert "transformer.blocks.0.attn1.to_out.0.weight" in keys and "transformer.blocks.0.ffn.net.0.proj.bias" in keys
    assert "transformer.condition_embedder.time_proj.weight" in keys and "transformer.scale_shift_table" in keys
    assert set(WO.make_state_dict(cfg, 0).keys()) | {"transformer.logvar_linear.weight", "transformer.logvar_linear.bias"} == keys
    # original-Wan naming (reference Wan/network.py:1012-1046)
    sd = WO.make_state_dict(cfg, 0)
    inv = [("condition_embedder.time_embedder.linear_1", "time_embedding.0"), ("condition_embedder.time_embedder.linear_2", "time_embedding.2"),
           ("condition_embedder.text_embedder.linear_1", "text_embedding.0"), ("condition_embedder.text_embedder.linear_2", "text_embedding.2"),
           ("condition_embedder.time_proj", "time_projection.1"), ("proj_out", "head.head"),
           ("ffn.net.0.proj", "ffn.0"), ("ffn.net.2", "ffn.2"), ("norm2", "norm3"),
           ("attn1.to_q", "self_attn.q"), ("attn1.to_k", "self_attn.k"), ("attn1.to_v", "self_attn.v"), ("attn1.to_out.0", "self_attn.o"),
           ("attn1.norm_q", "self_attn.norm_q"), ("attn1.norm_k", "self_attn.norm_k"),
           ("attn2.to_q", "cross_attn.q"), ("attn2.to_k", "cross_attn.k"), ("attn2.to_v", "cross_attn.v"), ("attn2.to_out.0", "cross_attn.o"),
           ("attn2.norm_q", "cross_attn.norm_q"), ("attn2.norm_k", "cross_attn.norm_k")]
    orig = {}
    for k, v in sd.items():
        k = k[len("transformer."):]
        if k == "scale_shift_table":
            k = "head.modulation"
        elif k.endswith("scale_shift_table"):
            k = k.replace("scale_shift_table", "modulation")
        for a, b in inv:
            if a in k:
                k = k.replace(a, b)
                break
        orig["model." + k] = v
    missing, unexpected = net.load_state_dict({"generator": orig}, strict=False)
    assert not unexpected and all("logvar" in m for m in missing), (missing, unexpected)
    for k, v in sd.items():
        assert torch.equal(net.state_dict()[k], v), k


def test_wan_teacher_sampler_runs_on_the_gpu_and_tracks_the_oracle():
    """`Wan.sample` end to end on the GPU (the loop itself is pinned to the reference on the CPU,
    tests/test_wan_sample_cpu.py; the forward by tests/test_wan_parity_gpu.py): 4 UniPC steps with guidance 2 against the
    same loop driven by the fp32 oracle forward.  Also checks that two consecutive flow predictions do not alias."""
    import fastgen_b200.wan as W
    from fastgen_b200.unipc import FlowUniPCMultistepScheduler
    from oracle import wan_oracle as WO

    cfg = WO.tiny_config()
    sd = {k: v.bfloat16() for k, v in WO.make_state_dict(cfg, 0).items()}
    net = W.Wan(transformer_config=cfg, enable_logvar_linear=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(torch.bfloat16).cuda().eval()
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(2, 16, 3, 8, 12, generator=g).bfloat16()
    text, neg = (torch.randn(2, 24, cfg["text_dim"], generator=g).bfloat16() for _ in range(2))
    t = torch.tensor([0.8, 0.3], dtype=torch.float64, device="cuda")
    a = net(noise.cuda(), t, condition=text.cuda())
    a_copy = a.clone()
    b = net(noise.cuda(), t, condition=neg.cuda())
    assert a.data_ptr() != b.data_ptr() and torch.equal(a, a_copy)
    out = net.sample(noise.cuda(), condition=text.cuda(), neg_condition=neg.cuda(), guidance_scale=2.0, num_steps=4, shift=5.0)
    # the same loop on the CPU with the oracle forward, in fp32 arithmetic (the exact value) and in bf16 arithmetic (what the
    # reference network computes), both on the same bf16 weights.  Latents, flow predictions and the guidance combination
    # stay in the noise dtype, bf16, exactly as in the loop under test and in the reference: only the network arithmetic
    # differs.  Guidance 2 (flow = 2 cond - uncond) and the multistep corrector amplify the ~7e-3 per-forward rounding
    # error, so the bound is relative to the reference arithmetic's own deviation over the same trajectory.
    def oracle_loop(dtype):
        sch = FlowUniPCMultistepScheduler()
        sch.config.flow_shift = 5.0
        sch.set_timesteps(4)
        lat = (noise.double() * (sch.timesteps[0] / 1000).double()).bfloat16()
        for ts in sch.timesteps:
            # t reaches the network in the latent dtype, like in the reference loop (Wan/network.py:957-958): the rescaled
            # timestep 1000 t is then a bf16 number (spacing 4 near 1000), which the sinusoid embedding is sensitive to
            tt = (ts / 1000).expand(2).bfloat16()
            fc = WO.wan_forward(sd, cfg, lat, tt, text, fwd_pred_type="flow", dtype=dtype).bfloat16()
            fu = WO.wan_forward(sd, cfg, lat, tt, neg, fwd_pred_type="flow", dtype=dtype).bfloat16()
            lat = sch.step(fu + 2.0 * (fc - fu), ts, lat)[0]
        assert lat.dtype == torch.bfloat16
        return lat

    lat32, lat16 = oracle_loop(torch.float32), oracle_loop(torch.bfloat16)
    rel, rel_ref = _rel(out.float(), lat32.float()), _rel(lat16.float(), lat32.float())
    print(f"Wan.sample 4 UniPC steps, guidance 2: ours vs fp32 oracle {rel:.3e}, reference bf16 arithmetic {rel_ref:.3e}")
    assert out.dtype == torch.bfloat16 and rel < max(2e-2, 1.5 * rel_ref), (rel, rel_ref)
    # teacher-forced: every forward of the trajectory on the oracle's own latents stays within the per-step tolerance
    sch = FlowUniPCMultistepScheduler()
    sch.config.flow_shift = 5.0
    sch.set_timesteps(4)
    lat = (noise.double() * (sch.timesteps[0] / 1000).double()).bfloat16()
    for ts in sch.timesteps:
        tt = (ts / 1000).expand(2).bfloat16()
        fc = WO.wan_forward(sd, cfg, lat, tt, text, fwd_pred_type="flow", dtype=torch.float32).bfloat16()
        fu = WO.wan_forward(sd, cfg, lat, tt, neg, fwd_pred_type="flow", dtype=torch.float32).bfloat16()
        ours = net(lat.cuda(), tt.cuda(), condition=text.cuda())
        assert _rel(ours.float(), fc.float()) < TOL, _rel(ours.float(), fc.float())
        lat = sch.step(fu + 2.0 * (fc - fu), ts, lat)[0]


def test_inference_tensor_condition_is_accepted_and_never_served_stale():
    """prompt embeddings created under torch.inference_mode (the reference's video_model_inference.py) or
    by torch.stack of a list condition inside generator_fn are inference tensors, which have no version counter: the
    text-K/V cache must not read `_version`, and must recompute when such a tensor is updated in place."""
    cfg = WO.tiny_config()
    net, sd = _nets(cfg, 0, "Wan")
    x, text = _inputs(cfg, 2, 3, 8, 12)
    _, text2 = _inputs(cfg, 2, 3, 8, 12, seed=9)
    t = torch.tensor([0.8, 0.3], dtype=torch.float64)
    with torch.inference_mode():
        cond = text.cuda().clone()
        assert cond.is_inference()
        out1 = net(x.cuda(), t.cuda(), condition=cond, fwd_pred_type="x0")
        cond.copy_(text2.cuda())            # same storage, new contents, no version counter to reveal it
        out2 = net(x.cuda(), t.cuda(), condition=cond, fwd_pred_type="x0")
        out3 = net(x.cuda(), t.cuda(), condition=list(text2.cuda()), fwd_pred_type="x0")   # list -> stack inside forward
    ref1 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type="x0", dtype=torch.float32)
    ref2 = WO.wan_forward(sd, cfg, x, t, text2, fwd_pred_type="x0", dtype=torch.float32)
    assert _rel(out1, ref1) < TOL and _rel(out2, ref2) < TOL and torch.equal(out2, out3)
    assert _rel(out2, ref1) > 5 * TOL


def test_causal_teacher_sampler_runs_on_the_gpu_and_tracks_the_oracle():
    """`CausalWan.sample` end to end on the GPU: KV-cached chunks, cache tags pos / neg, guidance, cache-fill passes --
    against the same loop driven by the fp32 oracle forward (the loop itself is pinned to the reference on the CPU,
    tests/test_wan_sample_cpu.py)."""
    from fastgen_b200.unipc import FlowUniPCMultistepScheduler

    cfg = WO.tiny_config(layers=2)
    total = 6
    net, sd = _nets(cfg, 2, "CausalWan", total_num_frames=total, chunk_size=3)
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(1, 16, total, 8, 8, generator=g).bfloat16()
    text, neg = (torch.randn(1, 24, cfg["text_dim"], generator=g).bfloat16() for _ in range(2))
    out = net.sample(noise.cuda().clone(), condition=text.cuda(), neg_condition=neg.cuda(), guidance_scale=2.0,
                     sample_steps=3, shift=5.0)
    assert out.shape == noise.shape and out.dtype == torch.bfloat16 and bool(torch.isfinite(out.float()).all())
    # oracle loop: same chunking, same scheduler, oracle forwards with an explicit KV cache per tag
    x = noise.clone()
    fs = (8 // cfg["patch"][1]) * (8 // cfg["patch"][2])
    caches = {tag: WO.KVCache(cfg, 1, total * fs, 24, torch.float32) for tag in ("pos", "neg")}
    fwd = dict(total_num_frames=total, chunk_size=3, is_ar=True, fwd_pred_type="flow", dtype=torch.float32)
    sch = FlowUniPCMultistepScheduler()
    for start in (0, 3):
        xc = x[:, :, start:start + 3]
        sch.config.flow_shift = 5.0
        sch.set_timesteps(3)
        for ts in sch.timesteps:
            tt = (ts / 1000).expand(1)
            f = {tag: WO.causal_wan_forward(sd, cfg, xc, tt, c, cache=caches[tag], cur_start_frame=start, store_kv=False,
                                            **fwd).bfloat16() for tag, c in (("pos", text), ("neg", neg))}
            xc = sch.step(f["neg"] + 2.0 * (f["pos"] - f["neg"]), ts, xc)[0]
        x[:, :, start:start + 3] = xc
        t0 = torch.zeros(1, dtype=torch.bfloat16)
        for tag, c in (("pos", text), ("neg", neg)):
            WO.causal_wan_forward(sd, cfg, xc, t0, c, cache=caches[tag], cur_start_frame=start, store_kv=True, **fwd)
    rel = _rel(out.float(), x.float())
    assert rel < 5e-2, rel
