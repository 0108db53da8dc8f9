"""The vectorised sampler-step kernels (fg_edm_step / fg_axpby / fg_sub_div, csrc/sampler_vec.cuh) on hardware: equal to the
scalar forms bit for bit and to fp64 torch.  First run on a B200 in round 2 (all green, profiles/r2_logs/r2_unverified.log);
they are the default since."""
import pytest
import torch

pytestmark = pytest.mark.gpu


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
