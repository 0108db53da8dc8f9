"""Flow-matching UniPC multistep scheduler for the Wan teacher samplers (``Wan.sample`` / ``CausalWan.sample``).

The reference takes this object from a third-party dependency: ``diffusers.UniPCMultistepScheduler.from_pretrained(
model_id, subfolder="scheduler")`` (fastgen/networks/Wan/network.py:31, 635-639), which is neither under /root/reference
nor installed here.  RESTATED from the published diffusers 0.35.1 algorithm for exactly the configuration the Wan 2.1
scheduler files carry (``prediction_type="flow_prediction"``, ``use_flow_sigmas=True``, ``solver_order=2``,
``solver_type="bh2"``, ``predict_x0=True``, ``lower_order_final=True``, ``final_sigmas_type="zero"``, no thresholding, no
disabled corrector steps); PARITY UNPINNED against diffusers itself -- there is nothing to pin it to in this image.  The
sample loops only use the interface below (``config.flow_shift`` / ``config.num_train_timesteps``, ``set_timesteps``,
``timesteps``, ``step(...)[0]``), so a caller who has diffusers assigns the real object instead::

    net.unipc_scheduler = diffusers.UniPCMultistepScheduler.from_pretrained(model_id, subfolder="scheduler")

The arithmetic is scalar coefficients on the host plus a handful of elementwise passes over the latent per step (40-50
steps per teacher sample next to 2 network forwards each): plumbing, evaluated with torch in the latent's dtype exactly
as diffusers does.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Tuple

import numpy as np
import torch


class FlowUniPCMultistepScheduler:
    """UniPC (B(h) = expm1(h), data-prediction form) on the shifted flow-matching sigma grid."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, flow_shift: float = 3.0, solver_order: int = 2):
        if solver_order not in (1, 2, 3):
            raise ValueError("solver_order must be 1, 2 or 3")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, flow_shift=flow_shift,
                                      solver_order=solver_order, prediction_type="flow_prediction", use_flow_sigmas=True,
                                      solver_type="bh2", predict_x0=True, lower_order_final=True, final_sigmas_type="zero",
                                      disable_corrector=[])
        self.timesteps: Optional[torch.Tensor] = None
        self.sigmas: Optional[torch.Tensor] = None
        self.num_inference_steps = 0
        self._reset()

    def _reset(self) -> None:
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.config.solver_order
        self.timestep_list: List[Optional[torch.Tensor]] = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample: Optional[torch.Tensor] = None
        self.this_order = 1
        self._step_index: Optional[int] = None

    # ---- the sigma grid ---------------------------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        """sigma_i = shift * s / (1 + (shift - 1) s) on s = 1 - linspace(1, 1/T, N + 1), descending, the trailing 0 kept
        only as the last sigma; ``timesteps`` = sigma * T truncated to int64 (N entries)."""
        T, shift = self.config.num_train_timesteps, self.config.flow_shift
        alphas = np.linspace(1, 1 / T, num_inference_steps + 1)
        sigmas = 1.0 - alphas
        sigmas = np.flip(shift * sigmas / (1 + (shift - 1) * sigmas))[:-1].copy()
        timesteps = (sigmas * T).copy()
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)                     # stays on the host: only scalars are read from it
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self._reset()

    @property
    def step_index(self) -> Optional[int]:
        return self._step_index

    def _init_step_index(self, timestep) -> None:
        t = int(timestep) if not isinstance(timestep, torch.Tensor) else int(timestep.item())
        idx = (self.timesteps.cpu() == t).nonzero()
        if len(idx) == 0:
            self._step_index = len(self.timesteps) - 1
        else:
            # the first step of a schedule may repeat a timestep: take the second match, like diffusers
            self._step_index = int(idx[1 if len(idx) > 1 else 0].item())

    @staticmethod
    def _alpha_sigma(sigma: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return 1 - sigma, sigma

    def _lambda(self, sigma: torch.Tensor) -> torch.Tensor:
        a, s = self._alpha_sigma(sigma)
        return torch.log(a) - torch.log(s)

    # ---- one step ---------------------------------------------------------------------------------------------------
    def convert_model_output(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        """flow prediction -> data prediction: x0 = x - sigma v"""
        return sample - self.sigmas[self.step_index] * model_output

    def _coefficients(self, order: int, rks: List[torch.Tensor], h: torch.Tensor, device):
        rk = torch.tensor([float(r) for r in rks] + [1.0], device=device)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rk, i - 1))
            b.append(float(h_phi_k * fact / B_h))
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return torch.stack(R), torch.tensor(b, device=device), h_phi_1, B_h

    def _history(self, order: int, first_back: int, lambda_s0: torch.Tensor, h: torch.Tensor):
        m0 = self.model_outputs[-1]
        rks, d1s = [], []
        for i in range(1, order):
            si = self.step_index - (i + first_back)
            rk = (self._lambda(self.sigmas[si]) - lambda_s0) / h
            rks.append(rk)
            d1s.append((self.model_outputs[-(i + 1)] - m0) / rk)
        return rks, (torch.stack(d1s, dim=1) if d1s else None)

    def _predict(self, sample: torch.Tensor, order: int) -> torch.Tensor:
        """UniP: x_{t} from x_{s0}, the latest data prediction m0 and (order - 1) older ones"""
        m0, x = self.model_outputs[-1], sample
        sigma_t, sigma_s0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t = 1 - sigma_t
        lambda_s0 = self._lambda(sigma_s0)
        h = self._lambda(sigma_t) - lambda_s0
        rks, D1s = self._history(order, 0, lambda_s0, h)
        R, b, h_phi_1, B_h = self._coefficients(order, rks, h, x.device)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        if D1s is not None:
            rhos_p = (torch.tensor([0.5], dtype=x.dtype, device=x.device) if order == 2
                      else torch.linalg.solve(R[:-1, :-1], b[:-1]).to(x.dtype))
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, D1s)
        else:
            pred_res = 0
        return (x_t_ - alpha_t * B_h * pred_res).to(x.dtype)

    def _correct(self, this_model_output: torch.Tensor, last_sample: torch.Tensor, order: int) -> torch.Tensor:
        """UniC: re-derive x_{t} from x_{s0} once the data prediction AT t is known"""
        m0, x, model_t = self.model_outputs[-1], last_sample, this_model_output
        sigma_t, sigma_s0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t = 1 - sigma_t
        lambda_s0 = self._lambda(sigma_s0)
        h = self._lambda(sigma_t) - lambda_s0
        rks, D1s = self._history(order, 1, lambda_s0, h)
        R, b, h_phi_1, B_h = self._coefficients(order, rks, h, x.device)
        rhos_c = (torch.tensor([0.5], dtype=x.dtype, device=x.device) if order == 1
                  else torch.linalg.solve(R, b).to(x.dtype))
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], D1s) if D1s is not None else 0
        return (x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * (model_t - m0))).to(x.dtype)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = False):
        if self.num_inference_steps == 0 or self.timesteps is None:
            raise ValueError("call set_timesteps before step")
        if self.step_index is None:
            self._init_step_index(timestep)
        use_corrector = self.step_index > 0 and self.last_sample is not None
        converted = self.convert_model_output(model_output, sample)
        if use_corrector:
            sample = self._correct(converted, self.last_sample, self.this_order)
        for i in range(self.config.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
            self.timestep_list[i] = self.timestep_list[i + 1]
        self.model_outputs[-1] = converted
        self.timestep_list[-1] = timestep
        this_order = min(self.config.solver_order, len(self.timesteps) - self.step_index)   # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)                            # multistep warm-up
        self.last_sample = sample
        prev_sample = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return (prev_sample,)
