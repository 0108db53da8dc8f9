"""SURVEY.md 8f-2 hand-off: the oracle's restatement of the reference's VAE / text-encoder wrapper arithmetic against the
fixture written by the UNMODIFIED reference classes (oracle/make_golden.py:handoff_fixture), and the host-side pieces of
fastgen_b200.encoders that need no GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import handoff_oracle as HO
from oracle import ref_handoff_harness as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff_reference.npz")


def _g(d, k, bf16=False):
    a = torch.from_numpy(np.array(d[k]))
    return a.view(torch.bfloat16) if bf16 else a


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_wan_decode_and_encode_arithmetic_bit_exact(gold):
    vae = H.ToyWanVAE().eval().requires_grad_(False)
    cfg = vae.config
    for name, bf in (("f32", False), ("bf16", True)):
        lat = _g(gold, f"wan_latent_{name}", bf)
        z = HO.wan_decode_latents(lat, cfg.latents_mean, cfg.latents_std, vae.dtype)
        assert z.dtype == torch.float32
        out = HO.clip_video(vae.decode(z, return_dict=False)[0])
        assert torch.equal(out, _g(gold, f"wan_decode_{name}"))
        assert float(out.max()) == 1.0 and float(out.min()) == -1.0          # the clip is active in the fixture
    video = _g(gold, "wan_video")
    enc = HO.wan_encode_latents(vae.encode(video, return_dict=False)[0].sample(), cfg.latents_mean, cfg.latents_std, video.dtype)
    assert torch.equal(enc, _g(gold, "wan_encode_sample"))
    vb = video.bfloat16().to(vae.dtype)     # a bf16 video is cast to the VAE's dtype first (Wan/network.py:498)
    enc = HO.wan_encode_latents(vae.encode(vb, return_dict=False)[0].mode(), cfg.latents_mean, cfg.latents_std, vb.dtype)
    assert torch.equal(enc, _g(gold, "wan_encode_argmax"))


def test_wan_prompt_padding_bit_exact(gold):
    from fastgen_b200.encoders import prompt_clean

    enc, tok = H.toy_umt5(), H.ToyTokenizer(97)
    prompts = [prompt_clean(p) for p in H.PROMPTS]
    assert prompts[0] == "a corgi surfing a wave & smiling"      # double-unescaped entity, collapsed white space
    t = tok(prompts, max_length=512, return_attention_mask=True)
    lens = t.attention_mask.gt(0).sum(dim=1).long()
    hidden = enc(t.input_ids, t.attention_mask).last_hidden_state
    for name, dt, bf in (("f32", torch.float32, False), ("bf16", torch.bfloat16, True)):
        out = HO.pad_prompt_embeds(hidden, lens, dt)
        ref = _g(gold, f"wan_text_{name}", bf)
        assert out.shape == (3, 512, 32) and torch.equal(out, ref)
        for i, n in enumerate(lens.tolist()):
            assert float(ref[i, n:].float().abs().max()) == 0.0 and float(ref[i, :n].float().abs().min()) > 0.0


def test_sdxl_scaling_bit_exact(gold):
    vae = H.ToySDVAE().eval().requires_grad_(False)
    sf = vae.config.scaling_factor
    lat = _g(gold, "sdxl_latent")
    out = HO.clip_video(vae.decode(HO.sdxl_scale_decode(lat, sf), return_dict=False)[0])
    assert torch.equal(out, _g(gold, "sdxl_decode_f32"))
    vb = H.ToySDVAE().eval().requires_grad_(False).bfloat16()
    out = HO.clip_video(vb.decode(HO.sdxl_scale_decode(lat.bfloat16(), sf), return_dict=False)[0])
    assert torch.equal(out, _g(gold, "sdxl_decode_bf16", True))
    img = _g(gold, "sdxl_image")
    assert torch.equal(HO.sdxl_scale_encode(vae.encode(img, return_dict=False)[0].sample(), sf), _g(gold, "sdxl_encode_f32"))


def test_sdxl_text_hand_off_matches_reference(gold):
    """SDXLTextEncoder.encode is pure glue around the two CLIP encoders: our mirror runs on the CPU as it is"""
    from fastgen_b200.encoders import SDXLTextEncoder

    one, two = H.toy_clip_pair()
    te = SDXLTextEncoder(one, two, H.ToyTokenizer(97), H.ToyTokenizer(97))
    for name, dt, bf in (("f32", torch.float32, False), ("bf16", torch.bfloat16, True)):
        pe, pooled = te.encode(H.PROMPTS, precision=dt)
        assert pe.dtype == dt and pe.shape == (3, 77, 80) and pooled.shape == (3, 40)
        assert torch.equal(pe, _g(gold, f"sdxl_text_{name}", bf)) and torch.equal(pooled, _g(gold, f"sdxl_pooled_{name}", bf))


def test_wrappers_fail_loudly_without_a_gpu():
    from fastgen_b200 import _lib
    from fastgen_b200.encoders import WanVideoEncoder

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ve = WanVideoEncoder(H.ToyWanVAE())
    with pytest.raises(_lib.FastGenB200Error):
        ve.decode(torch.zeros(1, 16, 1, 2, 2))
