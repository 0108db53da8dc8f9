"""SURVEY.md 8f-2 hand-off on the GPU: fg_latent_affine / fg_clamp / fg_zero_tail_rows against the oracle (bit-exact), and
fastgen_b200.encoders around the stand-in networks against the fixture the UNMODIFIED reference classes wrote."""
import os

import numpy as np
import pytest
import torch

from oracle import handoff_oracle as HO
from oracle import ref_handoff_harness as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff_reference.npz")


def _g(d, k, bf16=False):
    a = torch.from_numpy(np.array(d[k]))
    return a.view(torch.bfloat16) if bf16 else a


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 16, 3, 6, 10), (1, 16, 21, 60, 104), (3, 16, 1, 5, 7)])
def test_latent_affine_modes_bit_exact(dtype, shape):
    from fastgen_b200 import ops

    torch.manual_seed(5)
    x = (torch.randn(shape) * 1.7).to(dtype)
    mean, inv_std = HO.wan_coefficients(H.WAN_LATENTS_MEAN, H.WAN_LATENTS_STD, 16, dtype)
    a, b = inv_std.flatten().float().cuda(), mean.flatten().float().cuda()
    dec = ops.latent_affine(x.cuda(), a, b, 0, torch.float32)
    assert torch.equal(dec.cpu(), HO.wan_decode_latents(x, H.WAN_LATENTS_MEAN, H.WAN_LATENTS_STD, torch.float32))
    enc = ops.latent_affine(x.cuda(), a, b, 1, dtype)
    assert torch.equal(enc.cpu(), HO.wan_encode_latents(x, H.WAN_LATENTS_MEAN, H.WAN_LATENTS_STD, dtype))
    s = torch.tensor([1 / H.SDXL_SCALING_FACTOR], dtype=torch.float64).float().cuda()
    sc = ops.latent_affine(x.cuda(), s, None, 2, dtype)
    assert torch.equal(sc.cpu(), HO.sdxl_scale_decode(x, H.SDXL_SCALING_FACTOR))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_clamp_in_place_bit_exact(dtype):
    from fastgen_b200 import ops

    torch.manual_seed(6)
    x = (torch.randn(3, 3, 5, 33, 17) * 1.5).to(dtype)
    x[0, 0, 0, 0, 0] = float("nan")
    y = ops.clamp_(x.cuda().clone(), -1.0, 1.0).cpu()
    ref = HO.clip_video(x)
    assert torch.equal(torch.nan_to_num(y.float(), nan=7.0), torch.nan_to_num(ref.float(), nan=7.0))


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_zero_tail_rows_bit_exact(out_dtype):
    from fastgen_b200 import ops

    torch.manual_seed(7)
    h = torch.randn(4, 512, 96)
    lens = torch.tensor([1, 17, 512, 0])
    out = ops.zero_tail_rows(h.cuda(), lens.cuda(), out_dtype).cpu()
    assert torch.equal(out, HO.pad_prompt_embeds(h, lens, out_dtype))


def test_wan_video_encoder_matches_reference_wrapper():
    from fastgen_b200.encoders import WanVideoEncoder

    gold = np.load(GOLD)
    ve = WanVideoEncoder(H.ToyWanVAE()).to("cuda")
    for name, bf in (("f32", False), ("bf16", True)):
        out = ve.decode(_g(gold, f"wan_latent_{name}", bf).cuda())
        assert out.dtype == torch.float32 and torch.equal(out.cpu(), _g(gold, f"wan_decode_{name}"))
    video = _g(gold, "wan_video")
    assert torch.equal(ve.encode(video.cuda(), mode="sample").cpu(), _g(gold, "wan_encode_sample"))
    assert torch.equal(ve.encode(video.bfloat16().cuda(), mode="argmax").cpu(), _g(gold, "wan_encode_argmax"))
    with pytest.raises(ValueError):
        ve.encode(video.cuda(), mode="mean")


def test_sdxl_image_encoder_matches_reference_wrapper():
    from fastgen_b200.encoders import SDXLImageEncoder

    gold = np.load(GOLD)
    ie = SDXLImageEncoder(H.ToySDVAE()).to("cuda")
    assert torch.equal(ie.decode(_g(gold, "sdxl_latent").cuda()).cpu(), _g(gold, "sdxl_decode_f32"))
    assert torch.equal(ie.encode(_g(gold, "sdxl_image").cuda()).cpu(), _g(gold, "sdxl_encode_f32"))
    ib = SDXLImageEncoder(H.ToySDVAE().bfloat16()).to("cuda")
    assert torch.equal(ib.decode(_g(gold, "sdxl_latent").bfloat16().cuda()).cpu(), _g(gold, "sdxl_decode_bf16", True))


def test_wan_text_encoder_matches_reference_wrapper():
    """the UMT5 encoder itself is library code evaluated on the GPU here and on the CPU in the fixture: its rows agree to
    fp32 round-off; the wrapper's part -- rows behind each prompt's last token -- is exactly zero"""
    from fastgen_b200.encoders import WanTextEncoder

    gold = np.load(GOLD)
    te = WanTextEncoder(H.toy_umt5(), H.ToyTokenizer(97)).to("cuda")
    lens = [8, 2, 10]
    for name, dt, bf in (("f32", torch.float32, False), ("bf16", torch.bfloat16, True)):
        out = te.encode(H.PROMPTS, precision=dt)
        ref = _g(gold, f"wan_text_{name}", bf)
        assert out.dtype == dt and out.shape == ref.shape and out.is_cuda
        assert torch.allclose(out.cpu().float(), ref.float(), rtol=0, atol=2e-2 if bf else 1e-5)
        for i, n in enumerate(lens):
            assert float(out[i, n:].float().abs().max()) == 0.0
    one = te.encode(H.PROMPTS[1])
    assert one.shape == (1, 512, 32) and torch.allclose(one.cpu(), _g(gold, "wan_text_single_f32"), rtol=0, atol=1e-5)
