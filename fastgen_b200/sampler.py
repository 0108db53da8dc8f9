"""Few-step student samplers -- mirrors of the reference ``methods/`` sampler API.

  FastGenModel.generator_fn / _student_sample_loop      fastgen/methods/model.py:315-420
  MeanFlowModel._student_sample_loop                    fastgen/methods/consistency_model/mean_flow.py:336-381
  CausVidModel._student_sample_loop                     fastgen/methods/distribution_matching/causvid.py:87-185

Same classmethod names, arguments, return values and exceptions, so callers of the reference
(``scripts/inference/*``, the FID script, the ``gen_rand`` partials) can swap the class.  Differences are all on
the host side and invisible to results:
  * the t list is kept on the host as well, so ``t_list[-1] == 0`` and ``t_next > 0`` are decided without a
    device->host synchronisation per step (reference: ``.item()`` at model.py:410 and ``if t_next > 0`` at :356);
  * for an :class:`fastgen_b200.edm.EDMPrecond` network the x0 update and the re-noising of a step are ONE kernel
    (``fg_edm_step``) reading the network's raw channels-last output; other networks go through the generic
    schedule kernels (``fg_axpby`` / ``fg_sub_div``);
  * random draws use ``torch.randn_like`` in the reference's order, so a fixed seed gives the reference's noise.
"""
from __future__ import annotations

from contextlib import ExitStack, contextmanager
from typing import Any, List, Optional

import torch

from . import ops
from .edm import EDMPrecond
from .noise_schedule import expand_like


@contextmanager
def inference_mode(*modules: torch.nn.Module, precision_amp: Optional[torch.dtype] = None, device_type: str = "cuda"):
    """eval() + torch.inference_mode() (+ autocast), restoring .training on exit (reference basic_utils.py:89-125)."""
    mods = [m for m in modules if isinstance(m, torch.nn.Module)]
    prev = [m.training for m in mods]
    try:
        for m in mods:
            m.eval()
        with torch.inference_mode(), torch.autocast(dtype=precision_amp, device_type=device_type,
                                                    enabled=precision_amp is not None):
            yield
    finally:
        for m, was in zip(mods, prev):
            m.train(was)


def _host_t_list(net, student_sample_steps: int, t_list: Optional[List[float]]) -> torch.Tensor:
    """float64 CPU t list: the schedule's default or the caller's list (reference model.py:403-410)."""
    sched = net.noise_scheduler
    if t_list is None:
        tl = sched.get_t_list(sample_steps=student_sample_steps, device=torch.device("cpu"))
    else:
        assert len(t_list) - 1 == student_sample_steps, (
            f"t_list length (excluding zero) != student_sample_steps: {len(t_list) - 1} != {student_sample_steps}")
        tl = torch.tensor(t_list, dtype=sched.t_precision)
    assert tl[-1].item() == 0, "t_list[-1] must be zero"
    return tl


class FastGenModel:
    """Sampler half of ``fastgen.methods.model.FastGenModel`` (training half is out of scope)."""

    @classmethod
    def _student_sample_loop(cls, net, x: torch.Tensor, t_list: torch.Tensor, condition: Any = None,
                             student_sample_type: str = "sde", **kwargs) -> torch.Tensor:
        t_host = t_list.detach().cpu()
        # validate every t that reaches a schedule call once, on the host (the reference asserts inside each call,
        # paying a device sync each time; the trailing 0 is never passed to the schedule)
        with net.noise_scheduler.validated_t(t_host[:-1]):
            return cls._sample_loop_validated(net, x, t_host, t_list.to(x.device), condition, student_sample_type)

    @classmethod
    def _sample_loop_validated(cls, net, x, t_host, t_dev, condition, student_sample_type):
        batch_size = x.shape[0]
        has_preserve_hook = hasattr(net, "preserve_conditioning")
        sched = net.noise_scheduler
        fused = isinstance(net, EDMPrecond) and net.net_pred_type == "x0" and not has_preserve_hook
        x_pred = x
        for i in range(len(t_host) - 1):
            t_batch = t_dev[i].expand(batch_size)
            renoise = bool(t_host[i + 1] > 0)
            if renoise and student_sample_type not in ("sde", "ode"):
                raise NotImplementedError(
                    f"student_sample_type must be one of 'sde', 'ode' but got {student_sample_type}")
            if fused:
                # network -> raw F (channels-last) ; then x0 = c_skip x + c_out F and x_next in one kernel
                plan, c_skip, c_out = net.raw_forward(x, t_batch, condition)
                x = x.contiguous()
                x_pred = torch.empty_like(x)
                if net.drop_precond in ("output", "both"):
                    c_skip = c_out = None
                if renoise:
                    t_next = t_dev[i + 1].to(torch.float64).expand(batch_size).contiguous()
                    x_next = torch.empty_like(x)
                    a_n, s_n = sched._alpha(t_next).contiguous(), sched._sigma(t_next).contiguous()
                    if student_sample_type == "sde":
                        eps = torch.randn_like(x_pred)
                        ops.edm_step(plan.f_out, plan.f_ld, x, x_pred, c_skip, c_out, x_next=x_next, eps=eps,
                                     alpha_next=a_n, sigma_next=s_n)
                    else:
                        t_cur = t_batch.to(torch.float64)
                        ops.edm_step(plan.f_out, plan.f_ld, x, x_pred, c_skip, c_out, x_next=x_next,
                                     alpha_next=a_n, sigma_next=s_n, ode=True,
                                     alpha_cur=sched._alpha(t_cur).contiguous(),
                                     sigma_cur=sched.non_zero_clamp(sched._sigma(t_cur)).contiguous())
                    x = x_next
                else:
                    ops.edm_step(plan.f_out, plan.f_ld, x, x_pred, c_skip, c_out)
                continue
            x_pred = net(x, t_batch, condition=condition, fwd_pred_type="x0")
            if has_preserve_hook:
                x_pred = net.preserve_conditioning(x_pred, condition)
            if renoise:
                t_next_batch = t_dev[i + 1].expand(batch_size)
                if student_sample_type == "sde":
                    eps_infer = torch.randn_like(x_pred)
                else:
                    eps_infer = sched.x0_to_eps(xt=x, x0=x_pred, t=t_batch)
                x = sched._forward_process(x_pred, eps_infer, t_next_batch)
                if has_preserve_hook:
                    x = net.preserve_conditioning(x, condition)
        return x_pred

    @classmethod
    def generator_fn(cls, net, noise: torch.Tensor, student_sample_steps: int = 1,
                     t_list: Optional[List[float]] = None, data: Optional[torch.Tensor] = None,
                     precision_amp: Optional[torch.dtype] = None, **kwargs) -> torch.Tensor:
        with inference_mode(net, precision_amp=precision_amp, device_type=noise.device.type):
            tl = _host_t_list(net, student_sample_steps, t_list)
            sched = net.noise_scheduler
            assert sched.is_t_valid(tl[0]), f"t_init must be in [{sched.min_t}, {sched.max_t}], but got {tl[0]}"
            latents = sched._latents(noise, tl[0].to(noise.device))
            if data is not None:
                latents = latents + data
            return cls._student_sample_loop(net, latents, t_list=tl, **kwargs).to(dtype=noise.dtype)


class MeanFlowModel(FastGenModel):
    """x <- x - dt * u(x, t, r) with the average velocity u (``fwd_pred_type='flow'``)."""

    @classmethod
    def _student_sample_loop(cls, net, x: torch.Tensor, t_list: torch.Tensor, condition: Any = None,
                             student_sample_type: str = "sde", **kwargs) -> torch.Tensor:
        t_host = t_list.detach().cpu()
        with net.noise_scheduler.validated_t(t_host):   # r = t_next (0 included) reaches the schedule in ode mode
            return cls._sample_loop_validated(net, x, t_host, t_list.to(x.device), condition, student_sample_type)

    @classmethod
    def _sample_loop_validated(cls, net, x, t_host, t_dev, condition, student_sample_type):
        batch_size = x.shape[0]
        sched = net.noise_scheduler
        for i in range(len(t_host) - 1):
            t_cur_batch = t_dev[i].expand(batch_size)
            t_next_batch = t_dev[i + 1].expand(batch_size)
            if student_sample_type == "sde":
                u = net(x, t=t_cur_batch, condition=condition, r=torch.zeros_like(t_next_batch), fwd_pred_type="flow")
                # reference: delta_t cast to x.dtype, product and difference evaluated in x.dtype
                delta = expand_like(t_dev[i], x).to(x.dtype)
                x = x - delta * u
                if bool(t_host[i + 1] > 0):
                    x = sched._forward_process(x, torch.randn_like(x), t_next_batch)
            elif student_sample_type == "ode":
                u = net(x, t=t_cur_batch, condition=condition, r=t_next_batch, fwd_pred_type="flow")
                delta = expand_like(t_dev[i] - t_dev[i + 1], x).to(x.dtype)
                x = x - delta * u
            else:
                raise NotImplementedError(
                    f"student_sample_type must be one of 'sde', 'ode' but got {student_sample_type}")
        return x


class CausVidModel(FastGenModel):
    """Chunk-by-chunk autoregressive sampling with a KV-cache fill pass per chunk (also SelfForcing / CausalKD)."""

    @classmethod
    def _student_sample_loop(cls, net, x: torch.Tensor, t_list: torch.Tensor, condition: Any = None,
                             student_sample_type: str = "sde", context_noise: Optional[float] = 0, **kwargs
                             ) -> torch.Tensor:
        t_host = t_list.detach().cpu()
        # every t a forward or a schedule call of the loop sees: the list (its trailing 0 is the cache-fill timestep) and
        # the context-noise level rounded to the latent dtype -- validated once here, not per call on the device
        seen = [t_host] + ([torch.tensor([context_noise], dtype=x.dtype)] if context_noise > 0 else [])
        with net.noise_scheduler.validated_t(*seen):
            return cls._sample_loop_validated(net, x, t_host, t_list.to(x.device), condition, student_sample_type,
                                              context_noise, **kwargs)

    @classmethod
    def _sample_loop_validated(cls, net, x, t_host, t_dev, condition, student_sample_type, context_noise, **kwargs):
        net.clear_caches()
        batch_size, num_frames = x.shape[0], x.shape[2]
        chunk_size = net.chunk_size
        num_chunks, remaining = num_frames // chunk_size, num_frames % chunk_size
        sched = net.noise_scheduler
        for i in range(max(1, num_chunks)):
            if num_chunks == 0:
                start, end = 0, remaining
            else:
                start = 0 if i == 0 else chunk_size * i + remaining   # the first chunk absorbs the remainder
                end = chunk_size * (i + 1) + remaining
            x_next = x[:, :, start:end, ...]
            for step in range(len(t_host) - 1):
                t_cur = t_dev[step].expand(batch_size)
                x_cur = x_next
                x_next = net(x_cur, t_cur, condition=condition, fwd_pred_type="x0", cache_tag="pos",
                             cur_start_frame=start, store_kv=False, is_ar=True, **kwargs)
                if bool(t_host[step + 1] > 0):
                    t_chunk_next = t_dev[step + 1].expand(batch_size)
                    if student_sample_type == "sde":
                        eps_infer = torch.randn_like(x_next)
                    elif student_sample_type == "ode":
                        eps_infer = sched.x0_to_eps(xt=x_cur.contiguous(), x0=x_next, t=t_cur)
                    else:
                        raise NotImplementedError(
                            f"student_sample_type must be one of 'sde', 'ode' but got {student_sample_type}")
                    x_next = sched._forward_process(x_next, eps_infer, t_chunk_next)
            x[:, :, start:end, ...] = x_next
            x_cache, t_cache = x_next, t_dev[-1].expand(batch_size)
            if context_noise > 0:
                t_cache = torch.full((batch_size,), context_noise, device=x.device, dtype=x.dtype)
                x_cache = sched.forward_process(x_next, torch.randn_like(x_next), t_cache)
            _ = net(x_cache, t_cache, condition=condition, fwd_pred_type="x0", cache_tag="pos", cur_start_frame=start,
                    store_kv=True, is_ar=True, **kwargs)
        net.clear_caches()
        return x

    @classmethod
    def generator_fn_extrapolation(cls, net, noise: torch.Tensor, condition: Any = None, *, num_segments: int,
                                   overlap_frames: int, student_sample_steps: int = 1,
                                   student_sample_type: str = "sde", t_list: Optional[List[float]] = None,
                                   precision_amp: Optional[torch.dtype] = None, context_noise: Optional[float] = 0,
                                   **kwargs) -> torch.Tensor:
        """Multi-segment autoregressive generation with a VAE bridge between segments -- reference
        ``CausVidModel.generator_fn_extrapolation`` (distribution_matching/causvid.py:187-397): same arguments, call
        order, random draws and float32 ``t_list``.  ``net.vae`` (``decode`` / ``encode``) is the caller's; the VAE
        itself is outside this path.  One deliberate difference: with ``t_list=None`` the reference
        calls ``get_t_list(..., dtype=torch.float32)``, a keyword that method does not take (causvid.py:244 vs
        noise_schedule.py:259), so only its explicit-``t_list`` branch runs; here the default list is
        ``get_t_list(steps)`` cast to float32."""
        with inference_mode(net, precision_amp=precision_amp, device_type=noise.device.type), ExitStack() as trusted:
            if num_segments < 1:
                raise ValueError("num_segments must be >= 1")
            if overlap_frames > 0 and getattr(net, "vae", None) is None:
                raise ValueError("generator_fn_extrapolation requires a VAE instance via `vae` when overlap_frames > 0")
            batch_size, _, segment_frames, _, _ = noise.shape
            dtype, device = noise.dtype, noise.device
            chunk_size = net.chunk_size
            if segment_frames % chunk_size != 0:
                raise ValueError(f"Segment length {segment_frames} must be divisible by chunk_size {chunk_size}")
            if overlap_frames < 0 or overlap_frames >= segment_frames:
                raise ValueError("overlap_frames must be in [0, segment_frames)")
            if overlap_frames % chunk_size != 0:
                raise ValueError("overlap_frames must be divisible by chunk_size")
            sched = net.noise_scheduler
            if t_list is None:
                tl = sched.get_t_list(student_sample_steps, device=device).to(torch.float32)
            else:
                assert len(t_list) - 1 == student_sample_steps, (
                    f"t_list length (excluding zero) != student_sample_steps: {len(t_list) - 1} != {student_sample_steps}")
                tl = torch.tensor(t_list, device=device, dtype=torch.float32)
            t_host = tl.detach().cpu()
            assert t_host[-1].item() == 0, "t_list[-1] must be zero"
            seen = [t_host] + ([torch.tensor([context_noise], dtype=dtype)] if context_noise and context_noise > 0 else [])
            trusted.enter_context(sched.validated_t(*seen))   # one host-side check instead of a device sync per call
            fwd = dict(condition=condition, fwd_pred_type="x0", cache_tag="pos", is_ar=True)

            def prefill_caches(segment_latents: torch.Tensor, frames: int) -> None:
                start = 0
                t_zero = tl[-1].expand(batch_size)
                while start < frames:
                    end = min(start + chunk_size, frames)
                    net(segment_latents[:, :, start:end, ...], t_zero, cur_start_frame=start, store_kv=True, **fwd, **kwargs)
                    start = end

            def run_segment(segment_latents: torch.Tensor, prefill_frames: int) -> torch.Tensor:
                x = segment_latents.clone()
                net.clear_caches()
                if prefill_frames > 0:
                    prefill_caches(x, prefill_frames)
                    x[:, :, prefill_frames:, ...] = sched.latents(x[:, :, prefill_frames:, ...].contiguous(), t_init=tl[0])
                else:
                    x = sched.latents(x, t_init=tl[0])
                start = prefill_frames
                while start < segment_frames:
                    end = min(start + chunk_size, segment_frames)
                    x_next = x[:, :, start:end, ...]
                    for step in range(len(t_host) - 1):
                        t_cur = tl[step].expand(batch_size)
                        x_cur = x_next
                        x_next = net(x_cur, t_cur, cur_start_frame=start, store_kv=False, **fwd, **kwargs)
                        if bool(t_host[step + 1] > 0):
                            if student_sample_type == "sde":
                                eps_infer = torch.randn_like(x_next)
                            elif student_sample_type == "ode":
                                eps_infer = sched.x0_to_eps(xt=x_cur.contiguous(), x0=x_next, t=t_cur)
                            else:
                                raise NotImplementedError(
                                    f"student_sample_type must be one of 'sde', 'ode' but got {student_sample_type}")
                            x_next = sched.forward_process(x_next, eps_infer, tl[step + 1].expand(batch_size))
                    x[:, :, start:end, ...] = x_next
                    x_cache, t_cache = x_next, tl[-1].expand(batch_size)
                    if context_noise and context_noise > 0:
                        t_cache = torch.full((batch_size,), context_noise, device=device, dtype=dtype)
                        x_cache = sched.forward_process(x_next, torch.randn_like(x_next), t_cache)
                    net(x_cache, t_cache, cur_start_frame=start, store_kv=True, **fwd, **kwargs)
                    start = end
                net.clear_caches()
                return x

            segments = []
            current, prefill = noise, 0
            for idx in range(num_segments):
                seg = run_segment(current, prefill)
                segments.append(seg if idx == 0 or overlap_frames == 0 else seg[:, :, overlap_frames:, :, :])
                if idx == num_segments - 1:
                    break
                if overlap_frames == 0:
                    current, prefill = torch.randn_like(noise), 0
                    continue
                # bridge through pixels: decode, keep the last frames, encode back; all but the first overlapped latent
                # frame are reused directly
                decoded = net.vae.decode(seg)
                tail = net.vae.encode(decoded[:, :, -overlap_frames:, :, :]).to(dtype=dtype, device=device)
                if overlap_frames > 1:
                    tail = torch.cat([tail[:, :, :1, :, :], seg[:, :, -(overlap_frames - 1):, :, :]], dim=2)
                nxt = torch.randn_like(seg)
                nxt[:, :, :overlap_frames, :, :] = tail
                current, prefill = nxt, overlap_frames
            net.clear_caches()
            return torch.cat(segments, dim=2).to(dtype=noise.dtype)


SelfForcingModel = CausVidModel
