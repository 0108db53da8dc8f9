"""CPU restatement of the arithmetic the reference's VAE / text-encoder wrappers perform.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU legs).  Pinned: tests/test_oracle_golden.py reproduces
tests/golden/handoff_reference.npz -- written by the UNMODIFIED reference classes -- bit for bit with these functions.
Each function cites the reference lines it restates; eager torch on the CPU, so every op rounds to the tensor dtype as
the reference's does.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def wan_coefficients(latents_mean: Sequence[float], latents_std: Sequence[float], z_dim: int, dtype, device="cpu"):
    """(mean, 1 / std) as fastgen/networks/Wan/network.py:515-523 builds them: lists -> tensor -> cast -> reciprocal"""
    mean = torch.tensor(latents_mean).view(1, z_dim, 1, 1, 1).to(device, dtype)
    inv_std = 1.0 / torch.tensor(latents_std).view(1, z_dim, 1, 1, 1).to(device, dtype)
    return mean, inv_std


def wan_decode_latents(latent_images: torch.Tensor, latents_mean, latents_std, vae_dtype) -> torch.Tensor:
    """Wan/network.py:524-526: latents = latent_images / latents_std + latents_mean, then .to(vae.dtype)"""
    mean, inv_std = wan_coefficients(latents_mean, latents_std, latent_images.shape[1], latent_images.dtype,
                                     latent_images.device)
    return (latent_images / inv_std + mean).to(vae_dtype)


def wan_encode_latents(vae_latent: torch.Tensor, latents_mean, latents_std, image_dtype) -> torch.Tensor:
    """Wan/network.py:503-513: (latent_images - latents_mean) * latents_std with coefficients in the images' dtype"""
    mean, inv_std = wan_coefficients(latents_mean, latents_std, vae_latent.shape[1], image_dtype, vae_latent.device)
    return (vae_latent - mean) * inv_std


def clip_video(x: torch.Tensor) -> torch.Tensor:
    """Wan/network.py:527, SDXL/network.py:147: .clip_(-1.0, 1.0)"""
    return x.clone().clip_(-1.0, 1.0)


def pad_prompt_embeds(hidden: torch.Tensor, seq_lens: torch.Tensor, precision, max_length: int = 512) -> torch.Tensor:
    """Wan/network.py:463-470: cast, keep the first seq_len rows of each prompt, pad with zeros to max_length, stack"""
    h = hidden.to(dtype=precision)
    rows: List[torch.Tensor] = [u[:v] for u, v in zip(h, seq_lens)]
    return torch.stack([torch.cat([u, u.new_zeros(max_length - u.size(0), u.size(1))]) for u in rows], dim=0)


def sdxl_scale_encode(vae_latent: torch.Tensor, scaling_factor: float) -> torch.Tensor:
    """SDXL/network.py:143: latent_images * scaling_factor"""
    return vae_latent * scaling_factor


def sdxl_scale_decode(latent_images: torch.Tensor, scaling_factor: float) -> torch.Tensor:
    """SDXL/network.py:146: 1 / scaling_factor * latent_images"""
    return 1 / scaling_factor * latent_images
