"""The restated flow-matching UniPC scheduler (fastgen_b200/unipc.py; diffusers is absent, so it cannot be pinned to the
original): known-answer checks that do not need the original --
  * the sigma / timestep grid against the closed form and hand-computed values,
  * a straight-line flow (constant velocity) is integrated exactly by every step (predictor and corrector),
  * the probability-flow ODE of a Gaussian data distribution has a closed-form solution: the sampler converges to it, with
    the order-2 multistep clearly ahead of order 1 at the same number of steps,
  * the last step returns the data prediction itself (sigma = 0) without NaNs."""
import math

import pytest
import torch

from fastgen_b200.unipc import FlowUniPCMultistepScheduler


def test_sigma_grid_matches_the_closed_form():
    s = FlowUniPCMultistepScheduler(flow_shift=5.0)
    s.set_timesteps(4)
    raw = [0.999, 0.74925, 0.4995, 0.24975]                       # 1 - linspace(1, 1/1000, 5), descending, without 0
    want = [5 * r / (1 + 4 * r) for r in raw]
    assert torch.allclose(s.sigmas[:-1].double(), torch.tensor(want, dtype=torch.float64), atol=1e-7)
    assert s.sigmas[-1].item() == 0.0 and len(s.sigmas) == 5
    assert s.timesteps.tolist() == [int(w * 1000) for w in want] == [999, 937, 833, 624]
    s.config.flow_shift = 1.0
    s.set_timesteps(50)
    assert s.timesteps[0].item() == 999 and s.timesteps.dtype == torch.int64 and len(s.timesteps) == 50
    assert torch.all(s.sigmas[:-1] > s.sigmas[1:])


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_constant_velocity_is_integrated_exactly(order, dtype):
    torch.manual_seed(0)
    x0, eps = torch.randn(2, 4, 3, 5, 5, dtype=dtype), torch.randn(2, 4, 3, 5, 5, dtype=dtype)
    s = FlowUniPCMultistepScheduler(flow_shift=3.0, solver_order=order)
    s.set_timesteps(7)
    x = (1 - s.sigmas[0]) * x0 + s.sigmas[0] * eps
    for i, t in enumerate(s.timesteps):
        x = s.step(eps - x0, t, x)[0]
        sig = s.sigmas[i + 1]
        assert torch.allclose(x, ((1 - sig) * x0 + sig * eps).to(dtype), atol=3e-6 if dtype == torch.float64 else 3e-5), i
    assert x.dtype == dtype and torch.allclose(x, x0, atol=3e-6 if dtype == torch.float64 else 3e-5)


def _gaussian_error(order, steps, sd=0.5):
    """data ~ N(0, sd^2): posterior mean x0(x, s) = (1-s) sd^2 x / var(s), var(s) = (1-s)^2 sd^2 + s^2, velocity
    (x - x0)/s = x (s - (1-s) sd^2) / var(s); the probability-flow solution keeps x / sqrt(var(s)) constant.  Returns the
    relative error at the grid point in the middle of the schedule (raw sigma 0.4995: on the grid of every even N)."""
    sch = FlowUniPCMultistepScheduler(flow_shift=3.0, solver_order=order)
    sch.set_timesteps(steps)
    var = lambda s_: (1 - s_) ** 2 * sd * sd + s_ * s_  # noqa: E731
    x = torch.ones(1, 1, 1, 2, 2, dtype=torch.float64)
    for i, t in enumerate(sch.timesteps):
        sg = sch.sigmas[i].double().item()
        if i == steps // 2:
            exact = math.sqrt(var(sg) / var(sch.sigmas[0].double().item()))
            return abs(x.flatten()[0].item() - exact) / exact
        x = sch.step(x * (sg - (1 - sg) * sd * sd) / var(sg), t, x)[0]


def test_convergence_orders_on_the_gaussian_probability_flow():
    """UniPC-p with its corrector is of order p + 1: halving the step divides the error by ~4 (p = 1) and ~8 (p = 2)"""
    e1 = [_gaussian_error(1, n) for n in (10, 20, 40, 80)]
    e2 = [_gaussian_error(2, n) for n in (10, 20, 40, 80)]
    r1 = [e1[i] / e1[i + 1] for i in range(3)]
    r2 = [e2[i] / e2[i + 1] for i in range(3)]
    assert all(3.0 < r < 5.0 for r in r1), (e1, r1)
    assert all(6.0 < r < 11.0 for r in r2), (e2, r2)
    assert e2[3] < 2e-6 and e1[3] < 5e-5, (e1, e2)


def test_last_step_returns_the_data_prediction():
    s = FlowUniPCMultistepScheduler(flow_shift=5.0)
    s.set_timesteps(5)
    torch.manual_seed(2)
    x = torch.randn(2, 3, 1, 4, 4)
    for i, t in enumerate(s.timesteps):
        v = torch.randn_like(x)
        x0 = x - s.sigmas[i] * v
        x = s.step(v, t, x)[0]
        assert torch.isfinite(x).all()
    # the corrector may move the sample the last data prediction is formed from, so compare with the scheduler's own record
    assert torch.allclose(x, s.model_outputs[-1], atol=1e-6) and x0.shape == x.shape


def test_step_before_set_timesteps_raises_and_state_resets():
    s = FlowUniPCMultistepScheduler()
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), 999, torch.zeros(1))
    s.set_timesteps(3)
    x = torch.ones(1, 1, 1, 2, 2)
    for t in s.timesteps:
        x = s.step(torch.zeros_like(x), t, x)[0]
    assert s.step_index == 3 and torch.equal(x, torch.ones_like(x))
    s.set_timesteps(3)          # a second sample starts from a clean multistep history
    assert s.step_index is None and s.last_sample is None and s.lower_order_nums == 0
