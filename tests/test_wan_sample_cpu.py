"""Teacher sampling loops of the Wan networks (`Wan.sample`, `CausalWan.sample`) against traces and samples
of the UNMODIFIED reference loops (tests/golden/wan_sample_trace.npz, oracle/make_golden.py:wan_sample_fixture: the
reference's `Wan.sample` / `CausalWan.sample` run on a stub network with the reference's own rf schedule).

The loops are host logic around the network forward, so they are checked here on the CPU with the same stub forward:
call order, timestep values and dtypes, cache tags, chunk positions, store_kv flags, the classifier-free-guidance
combination and the in-place write-back must reproduce the reference BIT FOR BIT.  (The forwards themselves are covered
by the GPU parity tests; the schedule's two latent kernels are replaced by their torch formulas for this CPU run.)"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import wan_oracle as WO
from oracle.make_golden import wan_sample_stub_flow

GOLD = os.path.join(os.path.dirname(__file__), "golden", "wan_sample_trace.npz")


def _net(cls_name, **kw):
    import fastgen_b200.wan as W

    net = getattr(W, cls_name)(transformer_config=WO.tiny_config(layers=1), enable_logvar_linear=False, **kw).eval()
    ns = net.noise_scheduler
    # CPU stand-ins for the two latent-sized schedule kernels (fg_axpby): the reference's formulas, float64, one rounding
    bc = lambda c, x: c.reshape(-1, *([1] * (x.dim() - 1)))  # noqa: E731
    ns._latents = lambda noise, t: (noise.double() * bc(ns._sigma(t.double()), noise)).to(noise.dtype)
    ns._forward_process = lambda x, eps, t: (x.double() * bc(ns._alpha(t.double()), x)
                                             + eps.double() * bc(ns._sigma(t.double()), x)).to(x.dtype)
    trace, cache = [], {}

    def forward(x, t, condition=None, cache_tag=None, cur_start_frame=0, store_kv=False, is_ar=None, **kw2):
        extra = {k: v for k, v in kw2.items() if v not in (None, False) and v != {}}
        trace.append(["fwd", [float(v) for v in t.float().flatten()], str(t.dtype), float(condition.flatten()[0]), cache_tag,
                      int(cur_start_frame), bool(store_kv), is_ar, sorted(extra)])
        state = cache.get(cache_tag, 0.0)
        if store_kv:
            cache[cache_tag] = state + float(x.float().mean())
        return wan_sample_stub_flow(x, t, condition, cur_start_frame, store_kv, state)

    def clear():
        trace.append(["clear_caches"])
        cache.clear()

    net.forward = forward
    net.clear_caches = clear
    return net, trace


def _t(a, dtype=None):
    t = torch.from_numpy(np.asarray(a))
    if t.dtype == torch.int16:          # the fixtures store bfloat16 tensors as their bit patterns
        t = t.view(torch.bfloat16)
    return t.to(dtype) if dtype is not None else t


@pytest.mark.parametrize("tag,dtype", [("bidir_cfg", torch.float32), ("bidir_nocfg", torch.float32),
                                       ("bidir_bf16", torch.bfloat16)])
def test_bidirectional_teacher_sampler_reproduces_the_reference_loop(tag, dtype):
    g = np.load(GOLD)
    net, trace = _net("Wan")
    kw = json.loads(str(g[tag + "_kw"]))
    with torch.no_grad():
        out = net.sample(_t(g[tag + "_noise"], dtype), condition=_t(g["cond"], dtype), neg_condition=_t(g["neg"], dtype), **kw)
    ref_trace = [e for e in json.loads(str(g[tag + "_trace"])) if e[0] == "fwd"]
    assert trace == ref_trace
    assert len(trace) == kw["num_steps"] * (2 if kw["guidance_scale"] is not None else 1)
    assert out.dtype == dtype and torch.equal(out.float(), _t(g[tag + "_out"]).float())


@pytest.mark.parametrize("tag", ["causal_cfg", "causal_ctx", "causal_short"])
def test_causal_teacher_sampler_reproduces_the_reference_loop(tag):
    g = np.load(GOLD)
    net, trace = _net("CausalWan", chunk_size=3, total_num_frames=21)
    kw = json.loads(str(g[tag + "_kw"]))
    eps = [_t(g[f"{tag}_eps{i}"]) for i in range(64) if f"{tag}_eps{i}" in g]
    draws = iter(eps)
    orig = torch.randn_like
    torch.randn_like = lambda x, **k: next(draws).to(x.dtype)
    net.train()                                   # the loop must switch to eval and restore the mode afterwards
    try:
        noise = _t(g[tag + "_noise"])
        with torch.no_grad():
            out = net.sample(noise, condition=_t(g["cond"]), neg_condition=_t(g["neg"]), **kw)
    finally:
        torch.randn_like = orig
    assert net.training
    ref_trace = [e for e in json.loads(str(g[tag + "_trace"])) if e[0] in ("fwd", "clear_caches")]
    assert trace == ref_trace
    assert trace[0] == ["clear_caches"] and trace[-1] == ["clear_caches"]
    assert out is noise                                                        # chunks are written back in place
    assert torch.equal(out, _t(g[tag + "_out"]))
    assert next(draws, None) is None                                           # every recorded draw was consumed
    if tag == "causal_cfg":      # 7 frames, chunk 3: chunk 0 = frames 0..3 (remainder absorbed), chunk 1 = frames 4..6
        starts = sorted({e[5] for e in trace if e[0] == "fwd"})
        assert starts == [0, 4]
        fills = [e for e in trace if e[0] == "fwd" and e[6]]
        assert [e[4] for e in fills] == ["pos", "neg", "pos", "neg"] and all(e[1] == [0.0, 0.0] for e in fills)


def test_scheduler_can_be_replaced_and_other_schedules_are_refused():
    net, _ = _net("Wan")
    from fastgen_b200.unipc import FlowUniPCMultistepScheduler

    assert isinstance(net.unipc_scheduler, FlowUniPCMultistepScheduler)
    mine = FlowUniPCMultistepScheduler(solver_order=1)
    net.unipc_scheduler = mine                  # e.g. diffusers.UniPCMultistepScheduler.from_pretrained(...)
    assert net.unipc_scheduler is mine
    net.schedule_type = "edm"
    with pytest.raises(AssertionError, match="not supported"):
        net.sample(torch.zeros(1, 4, 3, 4, 4), condition=torch.zeros(1, 5, 8), neg_condition=torch.zeros(1, 5, 8))
