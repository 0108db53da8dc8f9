"""GPU tests of code paths written after round 1's GPU budget ended.  They have been checked on the CPU only (host run of
the same kernel source: tests/test_sampler_vec_host_cpu.py), so they are OPT-IN both in the product (default off) and
here: a kernel that has never run on hardware must not be able to poison the CUDA context of the regular `-m gpu` run.

    FASTGEN_B200_TEST_UNVERIFIED=1 python -m pytest tests/test_unverified_paths_gpu.py -m gpu -x -q

First thing to run in round 2; once green, make the vector forms the default and fold these into test_kernels_gpu.py.
(The other unverified path, the one-pass peer exchange, needs 2 GPUs: tests/test_wan_sp_gpu.py.)"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FASTGEN_B200_TEST_UNVERIFIED") != "1",
                                 reason="never run on hardware yet: set FASTGEN_B200_TEST_UNVERIFIED=1")]


@pytest.fixture()
def vec_on():
    from fastgen_b200 import ops

    prev = ops.sampler_vec(True)
    yield
    ops.sampler_vec(prev)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["last", "sde", "ode"])
@pytest.mark.parametrize("shape", [(5, 3, 16, 16), (256, 3, 64, 64), (3, 1, 2, 2)])
def test_edm_step_vector_form_equals_scalar_form(vec_on, dtype, mode, shape):
    from fastgen_b200 import ops

    torch.manual_seed(5)
    n, c, h, w = shape
    x_t = (torch.randn(n, c, h, w, device="cuda") * 3).to(dtype)
    eps = torch.randn(n, c, h, w, device="cuda").to(dtype)
    f = torch.randn(n, h, w, 4, device="cuda")
    t = torch.rand(n, device="cuda", dtype=torch.float64) * 80 + 0.01
    t_next = torch.rand(n, device="cuda", dtype=torch.float64) * 2 + 0.002
    ones = torch.ones_like(t)
    c_skip = 0.25 / (t**2 + 0.25)
    c_out = t * 0.5 / (t**2 + 0.25).sqrt()

    def run():
        x0, xn = torch.empty_like(x_t), torch.empty_like(x_t)
        if mode == "last":
            ops.edm_step(f, 4, x_t, x0, c_skip, c_out)
            return x0, None
        if mode == "sde":
            ops.edm_step(f, 4, x_t, x0, c_skip, c_out, x_next=xn, eps=eps, alpha_next=ones, sigma_next=t_next)
        else:
            ops.edm_step(f, 4, x_t, x0, c_skip, c_out, x_next=xn, alpha_next=ones, sigma_next=t_next, ode=True,
                         alpha_cur=ones, sigma_cur=t)
        return x0, xn

    v0, vn = run()
    ops.sampler_vec(False)
    s0, sn = run()
    ops.sampler_vec(True)
    torch.cuda.synchronize()
    assert torch.equal(v0, s0) and (vn is None or torch.equal(vn, sn))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
def test_axpby_sub_div_vector_forms_are_bit_exact(vec_on, dtype):
    from fastgen_b200 import ops

    torch.manual_seed(6)
    a = torch.randn(4, 16, 3, 10, 12, device="cuda").to(dtype)
    b = torch.randn_like(a)
    ca = torch.rand(4, device="cuda", dtype=torch.float64)
    cb = torch.rand(4, device="cuda", dtype=torch.float64) + 0.1
    bc = lambda v: v.reshape(4, 1, 1, 1, 1)  # noqa: E731
    assert torch.equal(ops.axpby(a, ca, b, cb), (a.double() * bc(ca) + b.double() * bc(cb)).to(dtype))
    assert torch.equal(ops.axpby(a, ca), (a.double() * bc(ca)).to(dtype))
    assert torch.equal(ops.sub_div(a, b, ca, cb), ((a.double() - b.double() * bc(ca)) / bc(cb)).to(dtype))


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
    # the same loop on the CPU with the oracle forward (fp32 arithmetic on the same bf16 weights)
    sch = FlowUniPCMultistepScheduler()
    sch.config.flow_shift = 5.0
    sch.set_timesteps(4)
    lat = (noise.double() * (sch.timesteps[0] / 1000).double()).float()
    for ts in sch.timesteps:
        tt = (ts / 1000).expand(2).bfloat16().double()
        fc = WO.wan_forward(sd, cfg, lat.bfloat16(), tt, text, fwd_pred_type="flow", dtype=torch.float32).float()
        fu = WO.wan_forward(sd, cfg, lat.bfloat16(), tt, neg, fwd_pred_type="flow", dtype=torch.float32).float()
        lat = sch.step(fu + 2.0 * (fc - fu), ts, lat)[0]
    rel = ((out.float().cpu().double() - lat.double()).norm() / lat.double().norm()).item()
    assert out.dtype == torch.bfloat16 and rel < 5e-2, rel
