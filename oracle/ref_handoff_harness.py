"""Stand-in networks for the reference's VAE / text-encoder wrappers.

TEST INFRASTRUCTURE ONLY.  The wrappers (fastgen/networks/Wan/network.py:431-539, SDXL/network.py:46-156) own only the
arithmetic around the external networks; oracle/make_golden.py:handoff_fixture() runs the UNMODIFIED wrapper classes (built
with ``__new__``: their constructors download checkpoints) around the objects below and stores what they return.  The
same objects, rebuilt from the same seeds, sit inside fastgen_b200.encoders in the tests, so any difference is the
wrapper arithmetic.  The text encoders are the real transformers classes at toy sizes (random weights); the VAEs are
deterministic functions with diffusers' call interface (diffusers is absent from this image).
"""
from __future__ import annotations

import types
from typing import List

import torch

WAN_LATENTS_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
                    -0.1922, -0.9497, 0.2503, -0.2921]
WAN_LATENTS_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
                   1.1253, 2.8251, 1.9160]
SDXL_SCALING_FACTOR = 0.13025


class _Dist:
    def __init__(self, z):
        self.z = z

    def sample(self):
        return self.z * 1.0

    def mode(self):
        return self.z * 0.5


class _ToyVAE(torch.nn.Module):
    """Deterministic stand-in with diffusers' VAE call interface.  Only elementwise eager ops and strided slicing (no
    reductions), so CPU and GPU evaluate it bit-identically and any difference in a test is the wrapper's arithmetic;
    the decoded values leave [-1, 1], so the wrapper's clip matters."""

    def __init__(self, z_dim: int, spatial_dims: int):
        super().__init__()
        self.gain = torch.nn.Parameter(0.3 + 0.1 * torch.arange(z_dim, dtype=torch.float32))
        self.z_dim, self.nd = z_dim, spatial_dims

    @property
    def device(self):
        return self.gain.device

    @property
    def dtype(self):
        return self.gain.dtype

    def encode(self, x, return_dict=False):
        sub = x[..., ::8, ::8]
        shape = [1, self.z_dim] + [1] * (x.dim() - 2)
        z = torch.cat([sub[:, i % 3:i % 3 + 1] for i in range(self.z_dim)], dim=1) * self.gain.view(shape)
        return (_Dist(z),)

    def decode(self, z, return_dict=False):
        v = z[:, 0:3] * 0.35 + z[:, 1:4] * 0.25
        v = v.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
        return (v,)


class ToyWanVAE(_ToyVAE):
    def __init__(self):
        super().__init__(16, 3)
        self.config = types.SimpleNamespace(latents_mean=WAN_LATENTS_MEAN, latents_std=WAN_LATENTS_STD, z_dim=16)


class ToySDVAE(_ToyVAE):
    def __init__(self):
        super().__init__(4, 2)
        self.config = types.SimpleNamespace(scaling_factor=SDXL_SCALING_FACTOR)


class ToyTokenizer:
    """Hugging Face tokenizer call interface over a character hash: ids in [3, vocab), 1 = end of sequence, 0 = padding"""

    def __init__(self, vocab: int, model_max_length: int = 77, eos: int = 1, pad: int = 0):
        self.vocab, self.model_max_length, self.eos, self.pad = vocab, model_max_length, eos, pad

    def __call__(self, texts, padding="max_length", max_length=None, truncation=True, add_special_tokens=True,
                 return_attention_mask=False, return_tensors="pt"):
        texts = [texts] if isinstance(texts, str) else list(texts)
        n = max_length or self.model_max_length
        ids = torch.full((len(texts), n), self.pad, dtype=torch.long)
        mask = torch.zeros(len(texts), n, dtype=torch.long)
        for i, t in enumerate(texts):
            toks: List[int] = [3 + (sum(ord(c) for c in w) * 31 + len(w)) % (self.vocab - 3) for w in t.split(" ") if w]
            toks = toks[: n - 1] + [self.eos]
            ids[i, : len(toks)] = torch.tensor(toks)
            mask[i, : len(toks)] = 1
        return types.SimpleNamespace(input_ids=ids, attention_mask=mask)


def toy_umt5(seed: int = 3):
    from transformers import UMT5Config, UMT5EncoderModel

    torch.manual_seed(seed)
    cfg = UMT5Config(vocab_size=97, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4, relative_attention_num_buckets=8,
                     relative_attention_max_distance=16, dropout_rate=0.0, feed_forward_proj="gated-gelu")
    return UMT5EncoderModel(cfg).eval().requires_grad_(False)


def toy_clip_pair(seed: int = 4):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection

    torch.manual_seed(seed)
    c1 = CLIPTextConfig(vocab_size=97, hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=4,
                        max_position_embeddings=77, projection_dim=24, eos_token_id=1, pad_token_id=0, bos_token_id=2)
    c2 = CLIPTextConfig(vocab_size=97, hidden_size=48, intermediate_size=96, num_hidden_layers=3, num_attention_heads=4,
                        max_position_embeddings=77, projection_dim=40, eos_token_id=1, pad_token_id=0, bos_token_id=2)
    return (CLIPTextModel(c1).eval().requires_grad_(False), CLIPTextModelWithProjection(c2).eval().requires_grad_(False))


PROMPTS = ["a  corgi surfing   a wave &amp;amp; smiling ", "sunset", "an astronaut riding a horse on mars, cinematic lighting"]
