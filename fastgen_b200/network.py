"""Network contracts of the drop-in boundary -- mirrors ``fastgen/networks/network.py:13-275``.

A B200 network is a ``torch.nn.Module`` with the reference's constructor keywords, attributes
(``net_pred_type``, ``schedule_type``, ``noise_scheduler``) and
``forward(x_t, t, condition, r, return_features_early, feature_indices, return_logvar, fwd_pred_type, **kw)``
signature, so the reference's ``FastGenModel.generator_fn`` and inference scripts can drive it unchanged.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Optional, Set

import torch

from .noise_schedule import NET_PRED_TYPES, get_noise_schedule


def tensor_identity(t: torch.Tensor):
    """Key under which results derived from a conditioning tensor (text keys / values) may be reused: storage pointer,
    version counter, shape, dtype.  ``None`` -- never reuse -- for tensors without a version counter: inference tensors
    (created under ``torch.inference_mode``, as the reference's ``video_model_inference.py`` does for its prompt
    embeddings, or by ``torch.stack`` of a list condition inside ``generator_fn``) cannot reveal an in-place update or a
    recycled allocation, so their derived values are recomputed on every call, like the reference does."""
    if t.is_inference():
        return None
    return (t.data_ptr(), t._version, tuple(t.shape), t.dtype)


class FastGenNetwork(ABC, torch.nn.Module):
    def __init__(self, net_pred_type: str = "x0", schedule_type: str = "edm", **net_kwargs):
        super().__init__()
        self.net_pred_type = net_pred_type
        self.schedule_type = schedule_type
        self._validate_net_pred_type(net_pred_type)
        self._noise_schedule_kwargs = dict(net_kwargs)
        self.set_noise_schedule(**net_kwargs)

    def _validate_net_pred_type(self, pred_type: str) -> None:
        if pred_type not in NET_PRED_TYPES:
            raise ValueError(f"Unsupported net_pred_type '{pred_type}'. Supported types are: {NET_PRED_TYPES}")

    def set_noise_schedule(self, schedule_type: Optional[str] = None, **noise_schedule_kwargs) -> None:
        if schedule_type is not None:
            self.schedule_type = schedule_type
        self.noise_scheduler = get_noise_schedule(self.schedule_type, **noise_schedule_kwargs)

    def reset_parameters(self) -> None:
        if getattr(self, "noise_scheduler", None) is not None:
            self.set_noise_schedule()

    def fully_shard(self, **kwargs):
        raise NotImplementedError(
            f"Network {self.__class__.__name__} is an inference network; FSDP sharding is a training-time feature."
        )

    def sample(self, noise: torch.Tensor, condition: Any = None, neg_condition: Any = None,
               guidance_scale: Optional[float] = 5.0, num_steps: int = 50, **kwargs) -> torch.Tensor:
        raise NotImplementedError(f"Network {self.__class__.__name__} does not implement the sample method.")

    @abstractmethod
    def forward(self, x_t: torch.Tensor, t: torch.Tensor, condition: Any = None, r: Optional[torch.Tensor] = None,
                return_features_early: bool = False, feature_indices: Optional[Set[int]] = None,
                return_logvar: bool = False, fwd_pred_type: Optional[str] = None, **fwd_kwargs):
        ...


class CausalFastGenNetwork(FastGenNetwork):
    """Adds the chunked autoregressive contract (reference network.py:211-275)."""

    def __init__(self, net_pred_type: str = "x0", schedule_type: str = "edm", chunk_size: int = 3,
                 total_num_frames: int = 21, **net_kwargs):
        super().__init__(net_pred_type=net_pred_type, schedule_type=schedule_type, **net_kwargs)
        self.chunk_size = chunk_size
        self.total_num_frames = total_num_frames

    @abstractmethod
    def clear_caches(self):
        ...
