"""Hand-off either side of the sampler: the reference's VAE / text-encoder wrappers.

Mirrors, with the same method names, arguments and results,

  * ``WanTextEncoder`` / ``WanVideoEncoder``    (fastgen/networks/Wan/network.py:431-539)
  * ``SDXLTextEncoder`` / ``SDXLImageEncoder``  (fastgen/networks/SDXL/network.py:46-156)

The networks inside them -- UMT5 / CLIP text encoders (transformers) and the Wan / SD VAEs (diffusers) -- are third-party
library code the reference only wraps; they stay external here too and are handed in as objects (``from_pretrained``
builds them like the reference does when the packages and weights are available).  What the REFERENCE wrote is the
arithmetic around them, and that runs as our kernels, one pass each instead of three to five eager ops:

  decode : latents / (1 / std) + mean, cast to the VAE's dtype   -> ``fg_latent_affine`` (mode 0)
           vae.decode(...)[0].clip_(-1, 1)                        -> ``fg_clamp`` in place
  encode : (vae latent - mean) * (1 / std)                        -> ``fg_latent_affine`` (mode 1)
  SDXL   : latent * scaling_factor,  1 / scaling_factor * latent  -> ``fg_latent_affine`` (mode 2)
  text   : u[:len] padded with zeros to max_length, cast           -> ``fg_zero_tail_rows``

Each op rounds like the eager op it replaces, so results are bit-identical to the reference's (tests/golden/
handoff_reference.npz, written by the UNMODIFIED reference classes around stand-in networks: oracle/make_golden.py).
"""
from __future__ import annotations

import html
import re
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops


def _basic_clean(text: str) -> str:
    """``basic_clean`` of the reference's prompt cleaning (fastgen/utils/basic_utils.py: ftfy.fix_text, then html.unescape
    twice, strip); ftfy is optional exactly like in the oracle's shims: without it the text passes through."""
    try:
        import ftfy

        text = ftfy.fix_text(text)
    except ImportError:
        pass
    return html.unescape(html.unescape(text)).strip()


def prompt_clean(text: str) -> str:
    """whitespace_clean(basic_clean(text)) -- what WanTextEncoder.encode applies to every prompt (Wan/network.py:454)"""
    return re.sub(r"\s+", " ", _basic_clean(text)).strip()


def _device_of(module: Any) -> torch.device:
    dev = getattr(module, "device", None)
    if dev is not None:
        return torch.device(dev)
    return next(module.parameters()).device


class WanTextEncoder:
    """fastgen/networks/Wan/network.py:431-477 around any encoder with the transformers ``UMT5EncoderModel`` call
    signature ``(input_ids, attention_mask) -> .last_hidden_state`` and any tokenizer with the Hugging Face call signature."""

    max_length = 512

    def __init__(self, text_encoder: Any, tokenizer: Any):
        self.text_encoder = text_encoder.eval().requires_grad_(False) if hasattr(text_encoder, "eval") else text_encoder
        self.tokenizer = tokenizer

    @classmethod
    def from_pretrained(cls, model_id_or_local_path: str, **kw) -> "WanTextEncoder":
        from transformers import AutoTokenizer, UMT5EncoderModel

        return cls(UMT5EncoderModel.from_pretrained(model_id_or_local_path, subfolder="text_encoder", **kw),
                   AutoTokenizer.from_pretrained(model_id_or_local_path, subfolder="tokenizer", **kw))

    def encode(self, conditioning: Optional[Any] = None, precision: torch.dtype = torch.float32) -> torch.Tensor:
        conditioning = [conditioning] if isinstance(conditioning, str) else conditioning
        conditioning = [prompt_clean(u) for u in conditioning]
        text_inputs = self.tokenizer(conditioning, padding="max_length", max_length=self.max_length, truncation=True,
                                     add_special_tokens=True, return_attention_mask=True, return_tensors="pt")
        ids, mask = text_inputs.input_ids, text_inputs.attention_mask
        dev = _device_of(self.text_encoder)
        seq_lens = mask.gt(0).sum(dim=1).long().to(dev)
        hidden = self.text_encoder(ids.to(dev), mask.to(dev)).last_hidden_state
        # reference: .to(precision), u[:v], cat with zeros to 512 rows, stack -- one pass here
        return ops.zero_tail_rows(hidden, seq_lens, precision)

    def to(self, *args, **kwargs) -> "WanTextEncoder":
        self.text_encoder.to(*args, **kwargs)
        return self


class WanVideoEncoder:
    """fastgen/networks/Wan/network.py:480-539 around any VAE with diffusers' ``AutoencoderKLWan`` interface:
    ``config.latents_mean / latents_std / z_dim``, ``encode(x, return_dict=False)[0].sample() | .mode()``,
    ``decode(z, return_dict=False)[0]``, ``device``, ``dtype``."""

    def __init__(self, vae: Any):
        self.vae = vae.eval().requires_grad_(False) if hasattr(vae, "eval") else vae
        self._coef: Dict[Tuple[torch.dtype, str], Tuple[torch.Tensor, torch.Tensor]] = {}

    @classmethod
    def from_pretrained(cls, model_id_or_local_path: str, **kw) -> "WanVideoEncoder":
        from diffusers.models import AutoencoderKLWan

        return cls(AutoencoderKLWan.from_pretrained(model_id_or_local_path, subfolder="vae", torch_dtype=torch.float32, **kw))

    def _coefficients(self, dtype: torch.dtype, device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
        """(1 / std, mean) rounded exactly like the reference builds them on every call -- ``torch.tensor(list)`` cast to
        the latent's dtype, then ``1.0 / std`` IN that dtype -- held as fp32 [C] for the kernel"""
        key = (dtype, str(device))
        if key not in self._coef:
            cfg = self.vae.config
            mean = torch.tensor(cfg.latents_mean).view(cfg.z_dim).to(device, dtype)
            inv_std = 1.0 / torch.tensor(cfg.latents_std).view(cfg.z_dim).to(device, dtype)
            self._coef[key] = (inv_std.float().contiguous(), mean.float().contiguous())
        return self._coef[key]

    def encode(self, real_images: torch.Tensor, mode: str = "sample") -> torch.Tensor:
        real_images = real_images.to(device=self.vae.device, dtype=self.vae.dtype)
        if mode == "sample":
            latent_images = self.vae.encode(real_images, return_dict=False)[0].sample()
        elif mode == "argmax":
            latent_images = self.vae.encode(real_images, return_dict=False)[0].mode()
        else:
            raise ValueError(f"Invalid mode: {mode}. Supported modes: ['sample', 'argmax']")
        inv_std, mean = self._coefficients(real_images.dtype, real_images.device)
        # the reference's arithmetic promotes (latent dtype, coefficient dtype): both are the VAE's dtype here
        return ops.latent_affine(latent_images.to(real_images.dtype), inv_std, mean, 1, real_images.dtype)

    def decode(self, latent_images: torch.Tensor) -> torch.Tensor:
        inv_std, mean = self._coefficients(latent_images.dtype, latent_images.device)
        latents = ops.latent_affine(latent_images, inv_std, mean, 0, self.vae.dtype)
        if latents.device != torch.device(self.vae.device):
            latents = latents.to(self.vae.device)
        videos = self.vae.decode(latents, return_dict=False)[0]
        return ops.clamp_(videos if videos.is_contiguous() else videos.contiguous(), -1.0, 1.0)

    def to(self, *args, **kwargs) -> "WanVideoEncoder":
        self.vae.to(*args, **kwargs)
        return self


class SDXLTextEncoder:
    """fastgen/networks/SDXL/network.py:46-128: two CLIP encoders, hidden_states[-2] of each concatenated on the feature
    axis, the pooled output of the second one."""

    def __init__(self, text_encoder_one: Any, text_encoder_two: Any, tokenizer_one: Any, tokenizer_two: Any):
        self.text_encoder_one = text_encoder_one.eval().requires_grad_(False)
        self.text_encoder_two = text_encoder_two.eval().requires_grad_(False)
        self.tokenizer_one, self.tokenizer_two = tokenizer_one, tokenizer_two

    @classmethod
    def from_pretrained(cls, model_id: str, **kw) -> "SDXLTextEncoder":
        from transformers import CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer

        return cls(CLIPTextModel.from_pretrained(model_id, subfolder="text_encoder", **kw),
                   CLIPTextModelWithProjection.from_pretrained(model_id, subfolder="text_encoder_2", **kw),
                   CLIPTokenizer.from_pretrained(model_id, subfolder="tokenizer", **kw),
                   CLIPTokenizer.from_pretrained(model_id, subfolder="tokenizer_2", **kw))

    def encode(self, conditioning: Optional[Any] = None, precision: torch.dtype = torch.float32) -> List[torch.Tensor]:
        ids = [tok(conditioning, padding="max_length", max_length=tok.model_max_length, truncation=True,
                   return_tensors="pt").input_ids for tok in (self.tokenizer_one, self.tokenizer_two)]
        embeds, pooled = [], None
        for text_input_ids, enc in zip(ids, (self.text_encoder_one, self.text_encoder_two)):
            out = enc(text_input_ids.to(_device_of(enc)), output_hidden_states=True)
            pooled = out[0].to(precision)
            h = out.hidden_states[-2].to(precision)
            embeds.append(h.view(h.shape[0], h.shape[1], -1))
        prompt_embeds = torch.cat(embeds, dim=-1)
        return [prompt_embeds, pooled.view(len(ids[0]), -1)]

    def to(self, *args, **kwargs) -> "SDXLTextEncoder":
        self.text_encoder_one.to(*args, **kwargs)
        self.text_encoder_two.to(*args, **kwargs)
        return self


class SDXLImageEncoder:
    """fastgen/networks/SDXL/network.py:131-156 around any VAE with diffusers' ``AutoencoderKL`` interface."""

    def __init__(self, vae: Any):
        self.vae = vae.eval().requires_grad_(False) if hasattr(vae, "eval") else vae

    @classmethod
    def from_pretrained(cls, model_id: str, **kw) -> "SDXLImageEncoder":
        from diffusers import AutoencoderKL

        return cls(AutoencoderKL.from_pretrained(model_id, subfolder="vae", **kw))

    @staticmethod
    def _scalar(v: float, device: torch.device) -> torch.Tensor:
        return torch.tensor([v], dtype=torch.float64).to(torch.float32).to(device)   # python scalar -> fp32, like torch's opmath

    def encode(self, real_images: torch.Tensor) -> torch.Tensor:
        latent_images = self.vae.encode(real_images, return_dict=False)[0].sample()
        s = self._scalar(self.vae.config.scaling_factor, latent_images.device)
        return ops.latent_affine(latent_images, s, None, 2, latent_images.dtype)

    def decode(self, latent_images: torch.Tensor) -> torch.Tensor:
        s = self._scalar(1 / self.vae.config.scaling_factor, latent_images.device)
        latents = ops.latent_affine(latent_images, s, None, 2, latent_images.dtype)
        images = self.vae.decode(latents, return_dict=False)[0]
        return ops.clamp_(images if images.is_contiguous() else images.contiguous(), -1.0, 1.0)

    def to(self, *args, **kwargs) -> "SDXLImageEncoder":
        self.vae.to(*args, **kwargs)
        return self
