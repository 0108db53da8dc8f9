"""Host-side mirror of the reference noise schedules (``fastgen/networks/noise_schedule.py``).

Same class / method names, argument meaning and error behaviour as the reference, restricted to what the
student sampler touches: ``get_t_list``, ``latents``, ``forward_process``,
``x0_to_eps``, ``eps_to_x0``, ``flow_to_x0``, ``x0_to_flow``, ``v_to_x0``, ``convert_model_output``,
``is_t_valid``, ``safe_clamp``, ``rescale_t`` and the scalar ``alpha/sigma`` functions.

Division of labour
  * time / coefficient arithmetic (a handful of float64 scalars) is torch plumbing, written so that ``t`` lists
    and index tensors are produced by the same torch ops as the reference and are bit-identical to it;
  * every pass over a latent tensor runs in the CUDA library (``fg_axpby`` / ``fg_sub_div``: float64 arithmetic,
    one rounding to the tensor dtype, like the reference's ``x.double()...to(dtype)``).  A CPU latent raises --
    there is deliberately no PyTorch fallback.

Training-time pieces of the reference schedules (``sample_t*``, trig/cogvideox schedules) are out of scope.
"""
from __future__ import annotations

from typing import Optional

from contextlib import contextmanager

import torch

NET_PRED_TYPES = {"x0", "eps", "v", "flow"}

_PRECISION = {"float64": torch.float64, "float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


def expand_like(x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Append singleton dims to ``x`` until it has ``target``'s rank (reference ``fastgen.utils.expand_like``)."""
    x = torch.atleast_1d(x)
    return x.reshape(tuple(x.shape) + (1,) * (target.dim() - x.dim())) if x.dim() < target.dim() else x


def _coef(c: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """fp64, contiguous, on the latent's device."""
    return c.to(device=like.device, dtype=torch.float64).contiguous()


class BaseNoiseSchedule(torch.nn.Module):
    """x_t = alpha(t) x_0 + sigma(t) eps   (reference noise_schedule.py:23-130)."""

    def __init__(self, min_t: float, max_t: float, num_steps: int, clamp_min: float = 1e-6,
                 t_precision: str = "float64", **kwargs):
        super().__init__()
        self._min_t = min_t
        self._max_t = max_t
        self.num_steps = num_steps
        self.clamp_min = clamp_min
        self.t_precision = _PRECISION[t_precision]
        self._trusted_depth = 0

    # ---- scalar properties -------------------------------------------------------------------------
    @property
    def min_t(self) -> float:
        assert self._min_t >= 0, "min_t must be non-negative"
        return self._min_t

    @property
    def max_t(self) -> float:
        assert self._max_t > self._min_t, "max_t must be greater than min_t"
        return self._max_t

    @property
    def max_sigma(self) -> float:
        raise NotImplementedError

    def _alpha(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def _sigma(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def _alpha_prime(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def _sigma_prime(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def _rescale_t(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def alpha(self, t: torch.Tensor) -> torch.Tensor:
        self._check_t(t, "t")
        return self._alpha(t)

    def sigma(self, t: torch.Tensor) -> torch.Tensor:
        self._check_t(t, "t")
        return self._sigma(t)

    def rescale_t(self, t: torch.Tensor) -> torch.Tensor:
        self._check_t(t, "t")
        return self._rescale_t(t)

    # ---- validity of t --------------------------------------------------------------------------------
    def _check_t(self, t: torch.Tensor, name: str = "t") -> None:
        """The reference's ``assert self.is_t_valid(t)`` at the top of every schedule call (:86, :436, ...).  On a device
        tensor that assert costs a device->host synchronisation per call; inside :meth:`validated_t` -- entered by the
        sampler loops of ``fastgen_b200.sampler`` after checking the whole t list once on the host -- it is skipped, so the
        host can queue all the forwards of a sample ahead of the GPU."""
        if self._trusted_depth > 0:
            return
        assert self.is_t_valid(t), f"{name} must be in [{self.min_t}, {self.max_t}], but got {t}"

    @contextmanager
    def validated_t(self, *t_host: torch.Tensor):
        """Validate the CPU tensors ``t_host`` (together: every t the enclosed calls will pass, each in the dtype it will
        be passed in -- the validity slack is one ulp of that dtype) once, then skip the per-call device-side checks inside
        the block.  Raises the same AssertionError as the per-call check."""
        for th in t_host:
            th = torch.as_tensor(th)
            assert not th.is_cuda, "validated_t takes the host copy of the t list"
            assert self.is_t_valid(th), f"t must be in [{self.min_t}, {self.max_t}], but got {th}"
        self._trusted_depth += 1
        try:
            yield
        finally:
            self._trusted_depth -= 1

    # ---- t helpers (reference :90-129, 259-272, 409-423) ---------------------------------------------
    @staticmethod
    def safe_clamp(t: torch.Tensor, min: Optional[float] = None, max: Optional[float] = None) -> torch.Tensor:
        """Clamp so that min <= t <= max also holds after rounding the bounds to ``t.dtype``."""
        def representable(bound: float, toward: float) -> float:
            b = torch.as_tensor(bound, dtype=t.dtype, device=t.device)
            v = b.item()
            if (toward > 0 and v < bound) or (toward < 0 and v > bound):
                v = torch.nextafter(b, torch.tensor(float("inf") * toward, dtype=t.dtype, device=t.device)).item()
            return v

        lo = representable(min, +1.0) if min is not None else None
        hi = representable(max, -1.0) if max is not None else None
        return torch.clamp(t, min=lo, max=hi)

    def non_zero_clamp(self, tensor: torch.Tensor) -> torch.Tensor:
        return torch.where(tensor >= 0, tensor.clamp(min=self.clamp_min), tensor.clamp(max=-self.clamp_min))

    def is_t_valid(self, t: torch.Tensor) -> torch.Tensor:
        dtype = t.dtype if t.dtype in (torch.bfloat16, torch.float32, torch.float64) else torch.float32
        lo = torch.as_tensor(self.min_t, dtype=dtype, device=t.device)
        hi = torch.as_tensor(self.max_t, dtype=dtype, device=t.device)
        lo = torch.nextafter(lo, torch.tensor(float("-inf"), dtype=dtype, device=t.device))
        hi = torch.nextafter(hi, torch.tensor(float("inf"), dtype=dtype, device=t.device))
        return torch.all((lo <= t) & (t <= hi))

    def get_t_list(self, sample_steps: int, device: Optional[torch.device] = None) -> torch.Tensor:
        t_list = torch.linspace(self.max_t, 0, sample_steps + 1, device=device, dtype=self.t_precision)
        return self.safe_clamp(t_list, max=self.max_t)

    # ---- latent-sized arithmetic: CUDA library ----------------------------------------------------------
    def latents(self, noise: torch.Tensor, t_init: Optional[torch.Tensor] = None) -> torch.Tensor:
        """noise * sigma(t_init), float64 arithmetic, result in noise.dtype (reference :72-88)."""
        from . import ops

        if t_init is None:
            t_init = torch.as_tensor(self.max_t, dtype=self.t_precision, device=noise.device)
        self._check_t(t_init, "t_init")
        return self._latents(noise, t_init)

    def _latents(self, noise: torch.Tensor, t_init: torch.Tensor) -> torch.Tensor:
        from . import ops

        sigma = _coef(self._sigma(t_init.to(torch.float64)).reshape(-1), noise)
        return ops.axpby(noise.contiguous(), sigma)

    def forward_process(self, x: torch.Tensor, eps: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """x*alpha(t) + eps*sigma(t) in float64, one rounding to x.dtype (reference :425-449)."""
        self._check_t(t, "t")
        return self._forward_process(x, eps, t)

    def _forward_process(self, x: torch.Tensor, eps: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        from . import ops

        t64 = t.to(torch.float64)
        return ops.axpby(x.contiguous(), _coef(self._alpha(t64).reshape(-1), x), eps.contiguous(),
                         _coef(self._sigma(t64).reshape(-1), x))

    def x0_to_eps(self, xt: torch.Tensor, x0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """(xt - x0*alpha) / clamp(sigma)  (reference :544-574)."""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        return ops.sub_div(xt.contiguous(), x0.contiguous(), _coef(self._alpha(t64).reshape(-1), xt),
                           _coef(self.non_zero_clamp(self._sigma(t64)).reshape(-1), xt))

    def eps_to_x0(self, xt: torch.Tensor, eps: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """(xt - eps*sigma) / clamp(alpha)  (reference :576-608)."""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        return ops.sub_div(xt.contiguous(), eps.contiguous(), _coef(self._sigma(t64).reshape(-1), xt),
                           _coef(self.non_zero_clamp(self._alpha(t64)).reshape(-1), xt))

    def flow_to_x0(self, xt: torch.Tensor, v: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """General form (reference :610-648): x0 = (xt*c - v) / clamp(c*alpha - alpha'),  c = sigma'/clamp(sigma).
        Evaluated as ((xt*c + v*(-1)) then / d) so each step is one fp64 op like the reference."""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        xt_coeff = self._sigma_prime(t64) / self.non_zero_clamp(self._sigma(t64))
        x0_coeff = self.non_zero_clamp(xt_coeff * self._alpha(t64) - self._alpha_prime(t64))
        num = ops.axpby(xt.double().contiguous(), _coef(xt_coeff.reshape(-1), xt), v.double().contiguous(),
                        _coef(-torch.ones_like(t64).reshape(-1), xt))
        zero = torch.zeros_like(num)
        out = ops.sub_div(num, zero, _coef(torch.zeros_like(t64).reshape(-1), xt), _coef(x0_coeff.reshape(-1), xt))
        return out.to(xt.dtype)

    def x0_to_flow(self, xt: torch.Tensor, x0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def v_to_x0(self, xt: torch.Tensor, v: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("Only AlphasNoiseSchedule supports v-prediction")

    def x0_to_v(self, xt: torch.Tensor, x0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("Only AlphasNoiseSchedule supports v-prediction")

    def convert_model_output(self, xt: torch.Tensor, model_output: torch.Tensor, t: torch.Tensor,
                             src_pred_type: str = "x0", target_pred_type: str = "eps") -> torch.Tensor:
        """src -> x0 -> target, same dispatch and exceptions as the reference (:666-726)."""
        if src_pred_type == target_pred_type:
            return model_output
        if src_pred_type not in NET_PRED_TYPES:
            raise ValueError(f"Unsupported src_pred_type '{src_pred_type}'. Supported types: {NET_PRED_TYPES}")
        if target_pred_type not in NET_PRED_TYPES:
            raise ValueError(f"Unsupported target_pred_type '{target_pred_type}'. Supported types: {NET_PRED_TYPES}")
        if src_pred_type == "v" or target_pred_type == "v":
            assert torch.allclose(self.alpha(t) ** 2 + self.sigma(t) ** 2, torch.ones_like(t), rtol=1e-2, atol=1e-2), \
                "Only AlphaNoiseSchedule supports v-prediction!"
        if src_pred_type == "x0":
            x0 = model_output
        elif src_pred_type == "eps":
            x0 = self.eps_to_x0(xt, model_output, t)
        elif src_pred_type == "v":
            x0 = self.v_to_x0(xt, model_output, t)
        else:
            x0 = self.flow_to_x0(xt, model_output, t)
        if target_pred_type == "x0":
            return x0
        if target_pred_type == "eps":
            return self.x0_to_eps(xt, x0, t)
        if target_pred_type == "v":
            return self.x0_to_v(xt, x0, t)
        return self.x0_to_flow(xt, x0, t)


class EDMNoiseSchedule(BaseNoiseSchedule):
    """Karras et al. 2022: x_t = x_0 + t*eps, t in [0.002, 80]  (reference :729-1034)."""

    def __init__(self, min_t: float = 0.002, max_t: float = 80.0, rho: float = 7.0, min_step_percent: float = 0.002,
                 max_step_percent: float = 0.998, num_steps: int = 1000, **kwargs):
        super().__init__(min_t, max_t, num_steps, **kwargs)
        # Karras rho-schedule, ascending in sigma (flipped like the reference so indices match bit for bit)
        ramp = torch.linspace(0, 1, num_steps, dtype=self.t_precision)
        lo, hi = min_t ** (1 / rho), max_t ** (1 / rho)
        self._sigmas = torch.flip((hi + ramp * (lo - hi)) ** rho, [0])
        self._min_step = int(min_step_percent * num_steps)
        self._max_step = int(max_step_percent * num_steps)

    def _rescale_t(self, t):
        return t

    @property
    def max_sigma(self) -> float:
        return self._max_t

    @property
    def sigmas(self) -> torch.Tensor:
        return self._sigmas

    def _alpha(self, t):
        return torch.ones_like(t)

    def _sigma(self, t):
        return t

    def _alpha_prime(self, t):
        return torch.zeros_like(t)

    def _sigma_prime(self, t):
        return torch.ones_like(t)

    def get_t_list_indices(self, sample_steps: int) -> torch.Tensor:
        """The int64 sigma-table indices behind ``get_t_list`` (bit-exact parity target)."""
        return torch.linspace(self._max_step, self._min_step, sample_steps + 1, device=self._sigmas.device).long()

    def get_t_list(self, sample_steps: int, device: Optional[torch.device] = None) -> torch.Tensor:
        t_list = self._sigmas[self.get_t_list_indices(sample_steps)].clone()
        t_list[-1] = 0.0
        t_list = t_list.to(device=device or self._sigmas.device, dtype=self.t_precision)
        return self.safe_clamp(t_list, max=self.max_t)

    def flow_to_x0(self, xt, v, t):
        """x0 = xt - v*t  (reference :975-1004)."""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        return ops.axpby(xt.contiguous(), _coef(torch.ones_like(t64).reshape(-1), xt), v.contiguous(),
                         _coef((-t64).reshape(-1), xt))

    def x0_to_flow(self, xt, x0, t):
        """(xt - x0) / clamp(t)  (reference :1006-1034)."""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        return ops.sub_div(xt.contiguous(), x0.contiguous(), _coef(torch.ones_like(t64).reshape(-1), xt),
                           _coef(self.non_zero_clamp(t64).reshape(-1), xt))


class RFNoiseSchedule(BaseNoiseSchedule):
    """Rectified flow: x_t = (1-t) x_0 + t eps  (reference :1306-1486)."""

    def __init__(self, min_t: float = 0.0, max_t: float = 0.999, num_steps: int = 1000, **kwargs):
        super().__init__(min_t, max_t, num_steps, **kwargs)
        assert 0 <= min_t < max_t <= 0.999, "RF min_t and max_t must be between 0 and 0.999"
        self._sigmas = torch.linspace(min_t, max_t, num_steps, dtype=self.t_precision)

    def _rescale_t(self, t):
        return self.num_steps * t

    @property
    def max_sigma(self) -> float:
        idx = int(self.num_steps * self.max_t)
        assert 0 <= idx < len(self._sigmas)
        return self._sigmas[idx].item()

    @property
    def sigmas(self):
        return self._sigmas

    def _alpha(self, t):
        return 1 - t

    def _sigma(self, t):
        return t

    def _alpha_prime(self, t):
        return -torch.ones_like(t)

    def _sigma_prime(self, t):
        return torch.ones_like(t)

    flow_to_x0 = EDMNoiseSchedule.flow_to_x0   # x0 = xt - v*t        (reference :1426-1455)
    x0_to_flow = EDMNoiseSchedule.x0_to_flow   # (xt - x0)/clamp(t)   (reference :1457-1486)


class AlphasNoiseSchedule(BaseNoiseSchedule):
    """DDPM-style table: alpha = sqrt(abar[int(1000 t)]), sigma = sqrt(1 - abar[...])  (reference :1036-1233)."""

    def __init__(self, alphas_cumprod: torch.Tensor, min_t: float = 0.0, max_t: float = 0.999, num_steps: int = 1000,
                 **kwargs):
        super().__init__(min_t, max_t, num_steps, **kwargs)
        assert 0 <= min_t < max_t <= 0.999, "Alphas min_t and max_t must be between 0 and 0.999"
        self._alphas_cumprod = alphas_cumprod.to(self.t_precision)
        assert len(self._alphas_cumprod) == num_steps, \
            f"alphas_cumprod's length {len(self._alphas_cumprod)} is not equal to num_steps {num_steps}"

    def _rescale_t(self, t):
        return self.num_steps * t

    @property
    def max_sigma(self) -> float:
        idx = int(self.num_steps * self.max_t)
        assert 0 <= idx < len(self._alphas_cumprod)
        return (1 - self._alphas_cumprod[idx]).sqrt().item()

    @property
    def sigmas(self):
        return (1 - self._alphas_cumprod).sqrt()

    def _t_to_idx(self, t: torch.Tensor) -> torch.Tensor:
        """integer table index: (1000*t).to(long) clamped to the table (bit-exact parity target)."""
        return torch.clamp(self._rescale_t(t).to(torch.long), 0, len(self._alphas_cumprod) - 1)

    def _alpha(self, t):
        return self._alphas_cumprod.to(t)[self._t_to_idx(t)].sqrt()

    def _sigma(self, t):
        return (1 - self._alphas_cumprod.to(t)[self._t_to_idx(t)]).sqrt()

    def _check_vp(self, t64: torch.Tensor, dtype: torch.dtype) -> None:
        tol = 1e-2 if dtype == torch.bfloat16 else 1e-3
        a, s = self._alpha(t64), self._sigma(t64)
        assert torch.allclose(a ** 2 + s ** 2, torch.ones_like(a), rtol=tol, atol=tol), \
            "v-prediction only supports the case of alpha_t^2+sigma_t^2 = 1."

    def v_to_x0(self, xt, v, t):
        """x0 = alpha*xt - sigma*v  (reference :1161-1202)."""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        self._check_vp(t64, xt.dtype)
        return ops.axpby(xt.contiguous(), _coef(self._alpha(t64).reshape(-1), xt), v.contiguous(),
                         _coef((-self._sigma(t64)).reshape(-1), xt))

    def x0_to_v(self, xt, x0, t):
        """v = (alpha*xt - x0) / clamp(sigma)  (reference :1204-1233), i.e. -(x0 - xt*alpha)/sigma."""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        self._check_vp(t64, xt.dtype)
        return ops.sub_div(x0.contiguous(), xt.contiguous(), _coef(self._alpha(t64).reshape(-1), xt),
                           _coef((-self.non_zero_clamp(self._sigma(t64))).reshape(-1), xt))

    def x0_to_flow(self, xt, x0, t):
        raise NotImplementedError("flow prediction is not defined for table schedules in the sampler path")


class TrigNoiseSchedule(BaseNoiseSchedule):
    """TrigFlow: x_t = cos(t) x_0 + sin(t) eps, t in [0, pi/2]  (reference :1489-1652; the sCM parameterisation)."""

    def __init__(self, min_t: float = 0.0, max_t: float = torch.pi / 2, num_steps: int = 1000, **kwargs):
        super().__init__(min_t, max_t, num_steps, **kwargs)
        assert 0 <= min_t < max_t, "Trig min_t must be non-negative and less than max_t"
        self._sigmas = torch.sin(torch.linspace(min_t, max_t, num_steps, dtype=self.t_precision))

    def _rescale_t(self, t):
        return t

    @property
    def max_sigma(self) -> float:
        return torch.sin(torch.tensor(self.max_t)).item()

    @property
    def sigmas(self):
        return self._sigmas

    def _alpha(self, t):
        return torch.cos(t)

    def _sigma(self, t):
        return torch.sin(t)

    def _alpha_prime(self, t):
        return -torch.sin(t)

    def _sigma_prime(self, t):
        return torch.cos(t)

    def sigma_idx_to_t(self, sigma_idx: torch.Tensor) -> torch.Tensor:
        assert sigma_idx.dtype == torch.long
        return sigma_idx.to(self.t_precision) / (self.num_steps - 1) * (self.max_t - self.min_t) + self.min_t

    def sqrt_snr(self, t: torch.Tensor) -> torch.Tensor:
        """cot(t) with tan(t) clamped away from zero (reference :1547-1557)"""
        self._check_t(t)
        return 1.0 / self.non_zero_clamp(torch.tan(t.to(torch.float64)))

    def sqrt_snr_to_t(self, sqrt_snr_t: torch.Tensor) -> torch.Tensor:
        x = sqrt_snr_t.to(torch.float64)
        return torch.atan2(torch.ones_like(x), x).to(sqrt_snr_t.dtype)

    def flow_to_x0(self, xt, v, t):
        """x0 = cos(t) x_t - sin(t) v in float64, one rounding (reference :1617-1634)"""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        return ops.axpby(xt.contiguous(), _coef(torch.cos(t64).reshape(-1), xt), v.contiguous(),
                         _coef(-torch.sin(t64).reshape(-1), xt))

    def x0_to_flow(self, xt, x0, t):
        """v = (cos(t) x_t - x_0) / clamp(sin(t)) in float64, one rounding (reference :1636-1652): evaluated as
        (x_0 - x_t cos(t)) / (-clamp(sin(t))), the same float64 operations up to two exact sign flips"""
        from . import ops

        self._check_t(t, "t")
        t64 = t.to(torch.float64)
        return ops.sub_div(x0.contiguous(), xt.contiguous(), _coef(torch.cos(t64).reshape(-1), xt),
                           _coef(-self.non_zero_clamp(torch.sin(t64)).reshape(-1), xt))


def scaled_linear_alphas_cumprod(num_steps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    """DDIM `scaled_linear` betas of SD/SDXL (diffusers scheduler_config: 0.00085 -> 0.012, 1000 steps), computed
    locally in float32 like diffusers' DDIMScheduler so that no hub download is needed."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_steps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class SDXLNoiseSchedule(AlphasNoiseSchedule):
    """reference :1255-1259 (DDIMScheduler.from_pretrained(...).alphas_cumprod), restated without the hub."""

    def __init__(self, *args, **kwargs):
        super().__init__(scaled_linear_alphas_cumprod(), *args, **kwargs)


SDNoiseSchedule = SDXLNoiseSchedule  # same scheduler constants for SD1.5 and SDXL

NOISE_SCHEDULES = {
    "edm": EDMNoiseSchedule,
    "alphas": AlphasNoiseSchedule,
    "sd": SDNoiseSchedule,
    "sdxl": SDXLNoiseSchedule,
    "rf": RFNoiseSchedule,
    "rectified_flow": RFNoiseSchedule,
    "trig": TrigNoiseSchedule,
}


def get_noise_schedule(name: str, **kwargs) -> BaseNoiseSchedule:
    if name not in NOISE_SCHEDULES:
        available = ", ".join(sorted(NOISE_SCHEDULES.keys()))
        raise KeyError(f"Unknown noise schedule '{name}'. Available schedules: {available}")
    return NOISE_SCHEDULES[name](**kwargs)
