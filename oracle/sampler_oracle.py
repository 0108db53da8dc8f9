"""CPU oracle for the sampler and the noise schedules (float64 torch / numpy arithmetic).

TEST INFRASTRUCTURE ONLY (see oracle/edm_oracle.py).  Restates, with the random draws passed in explicitly:
  fastgen/methods/model.py:315-420          FastGenModel._student_sample_loop / generator_fn
  fastgen/methods/consistency_model/mean_flow.py:336-381   MeanFlowModel._student_sample_loop
  fastgen/methods/distribution_matching/causvid.py:87-185  CausVidModel._student_sample_loop (call trace)
  fastgen/networks/noise_schedule.py:72-88, 259-272, 425-449, 544-608, 752-759, 940-1004, 1337-1341, 1426-1486
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def _bc(v: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    v = torch.atleast_1d(v)
    return v.reshape(tuple(v.shape) + (1,) * (like.dim() - v.dim()))


class Schedule:
    """alpha / sigma of the three schedules the sampler path uses, float64."""

    def __init__(self, kind: str, alphas_cumprod: Optional[torch.Tensor] = None):
        self.kind = kind
        if kind == "edm":
            self.min_t, self.max_t = 0.002, 80.0
            # ascending Karras rho = 7 table (:752-759); torch float64 ops so the table (and hence every t list)
            # is bit-identical to the reference's on the same host
            ramp = torch.linspace(0, 1, 1000, dtype=torch.float64)
            lo, hi = 0.002 ** (1 / 7.0), 80.0 ** (1 / 7.0)
            self.sigmas = torch.flip((hi + ramp * (lo - hi)) ** 7.0, [0]).numpy().copy()
        elif kind == "rf":
            self.min_t, self.max_t = 0.0, 0.999
        elif kind == "alphas":
            self.min_t, self.max_t = 0.0, 0.999
            self.abar = alphas_cumprod.double()
        else:
            raise KeyError(kind)

    def alpha(self, t: torch.Tensor) -> torch.Tensor:
        t = t.double()
        if self.kind == "edm":
            return torch.ones_like(t)
        if self.kind == "rf":
            return 1 - t
        return self.abar[(1000 * t).long().clamp(0, 999)].sqrt()

    def sigma(self, t: torch.Tensor) -> torch.Tensor:
        t = t.double()
        if self.kind in ("edm", "rf"):
            return t
        return (1 - self.abar[(1000 * t).long().clamp(0, 999)]).sqrt()

    def t_list_indices(self, steps: int) -> np.ndarray:
        """int64 sigma-table indices of the EDM t list: linspace(998, 2, steps+1) truncated (:957-962)."""
        return np.linspace(998.0, 2.0, steps + 1, dtype=np.float32).astype(np.int64)

    def get_t_list(self, steps: int) -> torch.Tensor:
        if self.kind == "edm":
            t = self.sigmas[self.t_list_indices(steps)].copy()
            t[-1] = 0.0
            return torch.from_numpy(t).clamp(max=self.max_t)
        return torch.linspace(self.max_t, 0, steps + 1, dtype=torch.float64).clamp(max=self.max_t)

    # latent arithmetic: float64, one rounding to the input dtype
    def latents(self, noise: torch.Tensor, t0: torch.Tensor) -> torch.Tensor:
        return (noise.double() * _bc(self.sigma(t0), noise)).to(noise.dtype)

    def forward_process(self, x: torch.Tensor, eps: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        return (x.double() * _bc(self.alpha(t), x) + eps.double() * _bc(self.sigma(t), x)).to(x.dtype)

    @staticmethod
    def _nz(v: torch.Tensor, c: float = 1e-6) -> torch.Tensor:
        return torch.where(v >= 0, v.clamp(min=c), v.clamp(max=-c))

    def x0_to_eps(self, xt, x0, t):
        return ((xt.double() - x0.double() * _bc(self.alpha(t), xt)) / _bc(self._nz(self.sigma(t)), xt)).to(xt.dtype)

    def eps_to_x0(self, xt, eps, t):
        return ((xt.double() - eps.double() * _bc(self.sigma(t), xt)) / _bc(self._nz(self.alpha(t)), xt)).to(xt.dtype)

    def flow_to_x0(self, xt, v, t):  # EDM and RF: x0 = xt - t v
        return (xt.double() - v.double() * _bc(t.double(), xt)).to(xt.dtype)

    def x0_to_flow(self, xt, x0, t):
        return ((xt.double() - x0.double()) / _bc(self._nz(t.double()), xt)).to(xt.dtype)


NetFn = Callable[..., torch.Tensor]  # net(x, t_batch, **kw) -> prediction


def student_sample_loop(net: NetFn, sched: Schedule, x: torch.Tensor, t_list: torch.Tensor,
                        eps_list: Sequence[torch.Tensor], sample_type: str = "sde",
                        trace: Optional[List[Tuple[str, torch.Tensor]]] = None) -> torch.Tensor:
    """model.py:315-372 with the `randn_like` draws supplied in call order."""
    b = x.shape[0]
    draws = iter(eps_list)
    x_pred = x
    for t_cur, t_next in zip(t_list[:-1], t_list[1:]):
        tb = t_cur.expand(b)
        x_pred = net(x, tb)
        if trace is not None:
            trace.append(("x0", x_pred))
        if t_next > 0:
            if sample_type == "sde":
                eps = next(draws)
            elif sample_type == "ode":
                eps = sched.x0_to_eps(x, x_pred, tb)
            else:
                raise NotImplementedError(f"student_sample_type must be one of 'sde', 'ode' but got {sample_type}")
            x = sched.forward_process(x_pred, eps, t_next.expand(b))
            if trace is not None:
                trace.append(("x_next", x))
    return x_pred


def generator_fn(net: NetFn, sched: Schedule, noise: torch.Tensor, steps: int = 1,
                 t_list: Optional[Sequence[float]] = None, eps_list: Sequence[torch.Tensor] = (),
                 sample_type: str = "sde", trace=None) -> torch.Tensor:
    """model.py:374-420."""
    if t_list is None:
        tl = sched.get_t_list(steps)
    else:
        assert len(t_list) - 1 == steps
        tl = torch.tensor(list(t_list), dtype=torch.float64)
    assert tl[-1].item() == 0, "t_list[-1] must be zero"
    lat = sched.latents(noise, tl[0])
    return student_sample_loop(net, sched, lat, tl, eps_list, sample_type, trace).to(noise.dtype)


def meanflow_sample_loop(net: NetFn, sched: Schedule, x: torch.Tensor, t_list: torch.Tensor,
                         eps_list: Sequence[torch.Tensor], sample_type: str = "sde") -> torch.Tensor:
    """mean_flow.py:336-381: net(x, t, r=...) returns the average velocity u (fwd_pred_type='flow')."""
    b = x.shape[0]
    draws = iter(eps_list)
    for t_cur, t_next in zip(t_list[:-1], t_list[1:]):
        tb, nb = t_cur.expand(b), t_next.expand(b)
        if sample_type == "sde":
            x = x - _bc(t_cur, x).to(x.dtype) * net(x, tb, r=torch.zeros_like(nb))
            if t_next > 0:
                x = sched.forward_process(x, next(draws), nb)
        elif sample_type == "ode":
            x = x - _bc(t_cur - t_next, x).to(x.dtype) * net(x, tb, r=nb)
        else:
            raise NotImplementedError(f"student_sample_type must be one of 'sde', 'ode' but got {sample_type}")
    return x


def causal_chunks(num_frames: int, chunk_size: int) -> List[Tuple[int, int]]:
    """causvid.py:117-129: remainder frames are absorbed by the first chunk."""
    nchunks, rem = num_frames // chunk_size, num_frames % chunk_size
    if nchunks == 0:
        return [(0, rem)]
    return [(0 if i == 0 else chunk_size * i + rem, chunk_size * (i + 1) + rem) for i in range(nchunks)]


def causvid_sample_loop(net: Callable[..., torch.Tensor], sched: Schedule, x: torch.Tensor, t_list: torch.Tensor,
                        chunk_size: int, eps_list: Sequence[torch.Tensor], sample_type: str = "sde",
                        context_noise: float = 0.0, clear_caches: Optional[Callable[[], None]] = None) -> torch.Tensor:
    """causvid.py:87-185 with the `randn_like` draws supplied in call order.
    ``net(x_chunk, t_batch, cur_start_frame=..., store_kv=...) -> x0`` (cache_tag 'pos', is_ar True implied)."""
    if clear_caches:
        clear_caches()
    b, frames = x.shape[0], x.shape[2]
    draws = iter(eps_list)
    x = x.clone()
    for start, end in causal_chunks(frames, chunk_size):
        x_next = x[:, :, start:end]
        for t_cur, t_next in zip(t_list[:-1], t_list[1:]):
            tb = t_cur.expand(b)
            x_cur = x_next
            x_next = net(x_cur, tb, cur_start_frame=start, store_kv=False)
            if t_next > 0:
                if sample_type == "sde":
                    eps = next(draws)
                elif sample_type == "ode":
                    eps = sched.x0_to_eps(x_cur, x_next, tb)
                else:
                    raise NotImplementedError(f"student_sample_type must be one of 'sde', 'ode' but got {sample_type}")
                x_next = sched.forward_process(x_next, eps, t_next.expand(b))
        x[:, :, start:end] = x_next
        x_cache, t_cache = x_next, t_list[-1].expand(b)
        if context_noise > 0:
            t_cache = torch.full((b,), context_noise, dtype=x.dtype)
            x_cache = sched.forward_process(x_next, next(draws), t_cache)
        net(x_cache, t_cache, cur_start_frame=start, store_kv=True)
    if clear_caches:
        clear_caches()
    return x


def causvid_extrapolation(net: Callable[..., torch.Tensor], sched: Schedule, noise: torch.Tensor, t_list: Sequence[float],
                          chunk_size: int, num_segments: int, overlap_frames: int, draws: Sequence[torch.Tensor],
                          sample_type: str = "sde", context_noise: float = 0.0,
                          vae_decode: Optional[Callable] = None, vae_encode: Optional[Callable] = None,
                          clear_caches: Optional[Callable[[], None]] = None) -> torch.Tensor:
    """causvid.py:187-397 with every ``randn_like`` draw supplied in call order; ``t_list`` is float32 there (:247-249).
    ``net(x_chunk, t_batch, cur_start_frame=..., store_kv=...) -> x0``."""
    b, frames = noise.shape[0], noise.shape[2]
    assert frames % chunk_size == 0 and 0 <= overlap_frames < frames and overlap_frames % chunk_size == 0
    tl = torch.tensor(list(t_list), dtype=torch.float32)
    assert tl[-1].item() == 0
    it = iter(draws)
    clear = clear_caches or (lambda: None)

    def run_segment(lat: torch.Tensor, prefill: int) -> torch.Tensor:
        x = lat.clone()
        clear()
        for s in range(0, prefill, chunk_size):
            net(x[:, :, s:min(s + chunk_size, prefill)], tl[-1].expand(b), cur_start_frame=s, store_kv=True)
        if prefill == 0:
            x = sched.latents(x, tl[0])
        else:
            x[:, :, prefill:] = sched.latents(x[:, :, prefill:], tl[0])
        for s in range(prefill, frames, chunk_size):
            e = min(s + chunk_size, frames)
            x_next = x[:, :, s:e]
            for t_cur, t_next in zip(tl[:-1], tl[1:]):
                tb = t_cur.expand(b)
                x_cur = x_next
                x_next = net(x_cur, tb, cur_start_frame=s, store_kv=False)
                if t_next > 0:
                    eps = next(it) if sample_type == "sde" else sched.x0_to_eps(x_cur, x_next, tb)
                    x_next = sched.forward_process(x_next, eps, t_next.expand(b))
            x[:, :, s:e] = x_next
            x_cache, t_cache = x_next, tl[-1].expand(b)
            if context_noise and context_noise > 0:
                t_cache = torch.full((b,), context_noise, dtype=noise.dtype)
                x_cache = sched.forward_process(x_next, next(it), t_cache)
            net(x_cache, t_cache, cur_start_frame=s, store_kv=True)
        clear()
        return x

    segs, cur, prefill = [], noise, 0
    for i in range(num_segments):
        seg = run_segment(cur, prefill)
        segs.append(seg if i == 0 or overlap_frames == 0 else seg[:, :, overlap_frames:])
        if i == num_segments - 1:
            break
        if overlap_frames == 0:
            cur, prefill = next(it), 0
            continue
        tail = vae_encode(vae_decode(seg)[:, :, -overlap_frames:]).to(noise.dtype)
        if overlap_frames > 1:
            tail = torch.cat([tail[:, :, :1], seg[:, :, -(overlap_frames - 1):]], dim=2)
        nxt = next(it).clone()
        nxt[:, :, :overlap_frames] = tail
        cur, prefill = nxt, overlap_frames
    clear()
    return torch.cat(segs, dim=2).to(noise.dtype)


def edm_teacher_sample(net: NetFn, sched: Schedule, noise: torch.Tensor, condition, neg_condition,
                       guidance_scale: Optional[float] = 5.0, num_steps: int = 50) -> torch.Tensor:
    """EDMPrecond.sample (EDM/network.py:976-1026): deterministic Euler steps on the EDM sigma list with optional CFG;
    ``net(x, t, condition=...) -> x0``; arithmetic in the tensor dtype like the reference."""
    sigmas = sched.get_t_list(num_steps)
    x = sched.latents(noise, sigmas[0])
    for sigma, sigma_next in zip(sigmas[:-1], sigmas[1:]):
        t = sigma.expand(x.shape[0])
        if guidance_scale is not None and guidance_scale > 1.0 and neg_condition is not None:
            x0 = net(torch.cat([x, x], 0), torch.cat([t, t], 0), condition=torch.cat([neg_condition, condition], 0))
            x0_uncond, x0_cond = x0.chunk(2)
            x0 = x0_uncond + guidance_scale * (x0_cond - x0_uncond)
        else:
            x0 = net(x, t, condition=condition)
        d = (x - x0) / _bc(t, x).to(x.dtype)
        x = x + (sigma_next - sigma).to(x.dtype) * d
    return x


def ddim_teacher_sample(net_eps: NetFn, sched: Schedule, noise: torch.Tensor, condition, neg_condition,
                        guidance_scale: Optional[float] = 7.5, num_steps: int = 50) -> torch.Tensor:
    """StableDiffusionXL.sample / StableDiffusion15.sample (SDXL/network.py:418-473): DDIM-style steps on the alphas
    schedule with optional CFG.  ``net_eps(x, t, condition) -> eps``; ``condition`` is a list of tensors that is batched
    [negative | positive] for guidance; arithmetic in the tensor dtype like the reference."""
    t_list = sched.get_t_list(num_steps)
    x = sched.latents(noise, t_list[0])
    b = x.shape[0]
    for t, t_next in zip(t_list[:-1], t_list[1:]):
        tb = t.expand(b)
        if guidance_scale is not None and guidance_scale > 1.0 and neg_condition is not None:
            cond = [torch.cat([n, c], 0) for n, c in zip(neg_condition, condition)]
            eps = net_eps(torch.cat([x, x], 0), torch.cat([tb, tb], 0), cond)
            eps_u, eps_c = eps.chunk(2)
            eps = eps_u + guidance_scale * (eps_c - eps_u)
        else:
            eps = net_eps(x, tb, condition)
        a, s_ = sched.alpha(tb), sched.sigma(tb)
        x0 = ((x.double() - eps.double() * _bc(s_, x)) / _bc(sched._nz(a), x)).to(x.dtype)    # eps_to_x0 (:576-608)
        an = sched.alpha(t_next.expand(b)).view(-1, 1, 1, 1).to(x.dtype)
        sn = sched.sigma(t_next.expand(b)).view(-1, 1, 1, 1).to(x.dtype)
        x = an * x0 + sn * eps
    return x
