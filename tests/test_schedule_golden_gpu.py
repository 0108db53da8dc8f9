"""The product's noise schedules (fastgen_b200/noise_schedule.py: one fg_axpby / fg_sub_div launch per call) against the
fixtures the UNMODIFIED reference schedules produced (tests/golden/schedules.npz, oracle/make_golden.py): bit-exact for
every latent-sized method of the EDM, RF (the Wan / MeanFlow schedule) and Alphas (SDXL) schedules, fp32 and bf16 latents,
per-sample and per-frame t.  (tests/test_oracle_golden.py checks the CPU oracle against the same file.)"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "schedules.npz")


def _t(a, bf16=False):
    t = torch.from_numpy(np.asarray(a))
    return (t.view(torch.bfloat16) if bf16 else t).cuda()


def _schedule(kind, g):
    from fastgen_b200.noise_schedule import AlphasNoiseSchedule, get_noise_schedule

    if kind == "alphas":
        return AlphasNoiseSchedule(torch.from_numpy(g["alphas_cumprod"]))
    return get_noise_schedule(kind)


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("kind", ["edm", "rf", "alphas"])
def test_schedule_latent_arithmetic_bit_exact_on_gpu(kind, dt):
    from fastgen_b200 import ops

    g = np.load(GOLD)
    bf = dt == "bf16"
    s = _schedule(kind, g)
    x, e, t = _t(g[f"x_{dt}"], bf), _t(g[f"e_{dt}"], bf), _t(g[f"{kind}_t3"])
    ops.launch_count_reset()
    assert torch.equal(s.latents(x, t[1]), _t(g[f"{kind}_latents_{dt}"], bf))
    assert torch.equal(s.forward_process(x, e, t), _t(g[f"{kind}_fwd_{dt}"], bf))
    assert torch.equal(s.x0_to_eps(x, e, t), _t(g[f"{kind}_x0_to_eps_{dt}"], bf))
    assert torch.equal(s.eps_to_x0(x, e, t), _t(g[f"{kind}_eps_to_x0_{dt}"], bf))
    n = 4
    if kind != "alphas":
        assert torch.equal(s.flow_to_x0(x, e, t), _t(g[f"{kind}_flow_to_x0_{dt}"], bf))
        assert torch.equal(s.x0_to_flow(x, e, t), _t(g[f"{kind}_x0_to_flow_{dt}"], bf))
        n = 6
    else:
        assert torch.equal(s.v_to_x0(x, e, t), _t(g[f"alphas_v_to_x0_{dt}"], bf))
        n = 5
    assert ops.launch_count() == n        # one kernel per schedule call: no eager casts or temporaries around it


def test_rf_convert_model_output_is_one_launch_and_matches_the_reference_override():
    """`convert_model_output(flow -> x0)` on the RF schedule -- what every Wan forward ends with -- is the reference's
    override `x0 = xt - v t` in float64 with one rounding (noise_schedule.py:1426-1455), as ONE launch."""
    from fastgen_b200 import ops

    g = np.load(GOLD)
    s = _schedule("rf", g)
    for dt in ("f32", "bf16"):
        bf = dt == "bf16"
        x, v, t = _t(g[f"x_{dt}"], bf), _t(g[f"e_{dt}"], bf), _t(g["rf_t3"])
        ops.launch_count_reset()
        with s.validated_t(t.cpu()):
            out = s.convert_model_output(x, v, t, src_pred_type="flow", target_pred_type="x0")
        assert ops.launch_count() == 1
        assert torch.equal(out, _t(g[f"rf_flow_to_x0_{dt}"], bf))
        back = s.convert_model_output(x, out, t, src_pred_type="x0", target_pred_type="flow")
        assert torch.equal(back, s.x0_to_flow(x, out, t))


def test_per_frame_forward_process_bit_exact_on_gpu():
    g = np.load(GOLD)
    s = _schedule("rf", g)
    out = s.forward_process(_t(g["video_x"]), _t(g["video_e"]), _t(g["video_t"])[:, None, :, None, None])
    assert torch.equal(out, _t(g["video_fwd"]))


def test_validated_t_skips_the_device_check_and_still_rejects_bad_lists():
    from fastgen_b200.noise_schedule import get_noise_schedule

    s = get_noise_schedule("rf")
    x = torch.randn(2, 4, 8, 8, device="cuda")
    bad = torch.tensor([1.5, 0.2], dtype=torch.float64, device="cuda")
    with pytest.raises(AssertionError):
        s.forward_process(x, x, bad)
    with pytest.raises(AssertionError):
        with s.validated_t(bad.cpu()):
            pass
    with s.validated_t(torch.tensor([0.5, 0.2], dtype=torch.float64)):
        s.forward_process(x, x, torch.tensor([0.5, 0.2], dtype=torch.float64, device="cuda"))
    assert s._trusted_depth == 0
