"""Parity at the BASELINE shapes ("next round" item 1): the pinned oracles evaluated ON THE GPU in true
float32 (oracle/device.py: TF32 off, explicit-matmul attention) as full-size checkers of the CUDA path.

  * fg_flash_attn at the Wan 1.3B production shapes: a 4 680-token chunk against a 28 080-key cache + itself (12 heads,
    with and without split-KV), the 32 760-token sequence bidirectional (ragged tail: 32 760 % 128 = 120) and block-causal;
  * Wan 1.3B (12 heads, 30 layers, 1536 wide): one bidirectional forward at 32 760 tokens, and the causal chunk sequence
    (7 chunks, each a denoising forward + the cache-fill forward) with per-chunk comparison;
  * SDXL-base configuration (2.57 G parameters), 1 sample at 128 x 128 latents, eps and x0;
  * EDM ImageNet-64 and CIFAR-10 networks at full width with compute="bf16x3" (the fp32 path): <= 1e-4.

Tolerances are north_star's: rel-L2 <= 1e-2 for the bf16 paths, <= 1e-4 for the fp32 path.  For the 30-layer Wan network
the reference's own bf16 arithmetic (the oracle evaluated in bf16) is measured against the same fp32 value, and the CUDA
path must not be further from it than that."""
import math

import pytest
import torch

from oracle import device as OD
from oracle import edm_oracle as EO
from oracle import sdxl_oracle as XO
from oracle import wan_oracle as WO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / b.norm()).item()


# ---------------------------------------------------------------------------------------------------------------------
# flash attention at production sizes
# ---------------------------------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, heads, d, mask=None):
    b, tq, _ = q.shape
    with OD.exact_fp32():
        o = OD.exact_sdpa(q.view(b, tq, heads, d).transpose(1, 2), k.reshape(b, -1, heads, d).transpose(1, 2),
                          v.reshape(b, -1, heads, d).transpose(1, 2), mask)
    return o.transpose(1, 2).flatten(2)


@pytest.mark.parametrize("l0", [28080, 14040, 0])
def test_flash_attn_wan_chunk_against_the_kv_cache(l0):
    """chunk 6 / chunk 3 / chunk 0 of a Wan 1.3B self-forcing sample: 4 680 queries, 12 heads of 128, keys = cache prefix +
    the chunk itself; with and without the split-KV workspace (148 SMs vs 12 x 19 query tiles)"""
    from fastgen_b200 import ops

    torch.manual_seed(11)
    b, tq, heads, d = 1, 4680, 12, 128
    dm = heads * d
    cap = 32760
    qkv = torch.randn(b, tq, 3 * dm, device=DEV).bfloat16()
    kc = torch.randn(b, cap, dm, device=DEV).bfloat16()
    vc = torch.randn(b, cap, dm, device=DEV).bfloat16()
    segs = ([(kc, vc, dm, cap * dm, l0)] if l0 else []) + [(qkv[..., dm:], qkv[..., 2 * dm:], 3 * dm, tq * 3 * dm, tq)]
    k_all = torch.cat([kc[:, :l0], qkv[..., dm:2 * dm]], 1)
    v_all = torch.cat([vc[:, :l0], qkv[..., 2 * dm:]], 1)
    ref = _attn_ref(qkv[..., :dm].contiguous(), k_all, v_all, heads, d)
    out = torch.empty(b, tq, dm, device=DEV, dtype=torch.bfloat16)
    ops.flash_attn(qkv, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    p = ops.flash_attn_params(qkv, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out, dm, tq * dm)
    ws = torch.empty(ops.flash_attn_workspace_size(p), device=DEV, dtype=torch.uint8)
    out2 = torch.zeros_like(out)
    ops.flash_attn(qkv, 3 * dm, tq * 3 * dm, segs, b, heads, d, tq, out2, dm, tq * dm, workspace=ws)
    torch.cuda.synchronize()
    # bf16 P and bf16 output: ~2^-9 per element; measured 2.3e-3 .. 2.6e-3
    assert _rel(out, ref) < 5e-3, _rel(out, ref)
    assert _rel(out2, ref) < 5e-3, _rel(out2, ref)


@pytest.mark.parametrize("causal", [False, True])
def test_flash_attn_full_480p_sequence(causal):
    """21 frames x 1560 tokens = 32 760 queries and keys, 12 heads: bidirectional, and the block-causal mask of the
    self-forcing teacher-forcing pass (chunks of 3 frames)"""
    from fastgen_b200 import ops

    torch.manual_seed(12)
    b, n, heads, d, fs, chunk = 1, 32760, 12, 128, 1560, 3
    dm = heads * d
    qkv = torch.randn(b, n, 3 * dm, device=DEV).bfloat16()
    mask = WO.block_causal_mask(21, fs, chunk).to(DEV) if causal else None
    ref = _attn_ref(qkv[..., :dm].contiguous(), qkv[..., dm:2 * dm], qkv[..., 2 * dm:], heads, d, mask)
    out = torch.empty(b, n, dm, device=DEV, dtype=torch.bfloat16)
    kw = dict(mask_first=chunk * fs, mask_chunk=chunk * fs) if causal else {}
    ops.flash_attn(qkv, 3 * dm, n * 3 * dm, [(qkv[..., dm:], qkv[..., 2 * dm:], 3 * dm, n * 3 * dm, n)], b, heads, d, n,
                   out, dm, n * dm, **kw)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 5e-3, _rel(out, ref)
    # row-wise as well: no query tile (the ragged last one included) may be off
    row = ((out.float() - ref.float()).norm(dim=-1) / ref.float().norm(dim=-1)).max().item()
    assert row < 2e-2, row


# ---------------------------------------------------------------------------------------------------------------------
# Wan 1.3B at full size
# ---------------------------------------------------------------------------------------------------------------------
def _wan_1_3b(cls_name, **kw):
    """the product network at the reference's 1.3B configuration with the reference's initialisers (+ a non-zero head),
    and its state dict (bf16, reference key names) for the oracle"""
    import fastgen_b200.wan as W

    torch.manual_seed(0)
    with torch.device(DEV):
        net = getattr(W, cls_name)(enable_logvar_linear=False, use_wan_official_sinusoidal=True, **kw)
        torch.nn.init.normal_(net.transformer.proj_out.weight, std=0.02)
    net = net.to(torch.bfloat16).eval()
    assert net.cfg["heads"] == 12 and net.cfg["layers"] == 30 and net.cfg["dim"] == 1536
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    return net, sd, dict(WO.WAN_1_3B)


def test_wan_1_3b_bidirectional_forward_at_32760_tokens():
    net, sd, cfg = _wan_1_3b("Wan")
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(1, 16, 21, 60, 104, device=DEV, generator=g).bfloat16()
    text = torch.randn(1, 512, 4096, device=DEV, generator=g).bfloat16()
    t = torch.tensor([0.937], dtype=torch.float64, device=DEV)
    with torch.no_grad(), OD.exact_fp32():
        ref32 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type="flow", dtype=torch.float32, official_sinusoidal=True)
        ref16 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type="flow", dtype=torch.bfloat16, official_sinusoidal=True)
        out = net(x, t, condition=text, fwd_pred_type="flow")
        x0 = net(x, t, condition=text, fwd_pred_type="x0")
        ref_x0 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type="x0", dtype=torch.float32, official_sinusoidal=True)
    torch.cuda.synchronize()
    e_ours, e_ref16 = _rel(out, ref32), _rel(ref16, ref32)
    print(f"wan 1.3B bidirectional 32760 tokens: flow rel-L2 ours {e_ours:.3e}, reference bf16 arithmetic {e_ref16:.3e}, "
          f"x0 {_rel(x0, ref_x0):.3e}")
    assert e_ours < max(1e-2, 1.25 * e_ref16), (e_ours, e_ref16)
    assert _rel(x0, ref_x0) < 1e-2, _rel(x0, ref_x0)


def test_wan_1_3b_causal_chunk_sequence_with_cache_fill():
    """all 7 chunks of a 21-frame sample: denoising forward (store_kv=False) then the cache-fill forward at t = 0
    (store_kv=True), product and oracle each building their own KV cache; compared per chunk (chunk 6 reads 28 080 cached
    keys written by 6 earlier cache-fill passes)"""
    net, sd, cfg = _wan_1_3b("CausalWan", total_num_frames=21, chunk_size=3)
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(1, 16, 21, 60, 104, device=DEV, generator=g).bfloat16()
    text = torch.randn(1, 512, 4096, device=DEV, generator=g).bfloat16()
    cache = WO.KVCache(cfg, 1, 32760, 512, torch.float32, device=DEV)
    fwd = dict(total_num_frames=21, chunk_size=3, is_ar=True, fwd_pred_type="x0", dtype=torch.float32,
               official_sinusoidal=True)
    errs = []
    net.clear_caches()
    with torch.no_grad(), OD.exact_fp32():
        for c in range(7):
            xc = x[:, :, 3 * c:3 * c + 3].contiguous()
            for tval, store in ((0.833, False), (0.0, True)):
                t = torch.full((1,), tval, dtype=torch.float64, device=DEV)
                out = net(xc, t, condition=text, fwd_pred_type="x0", cur_start_frame=3 * c, store_kv=store, is_ar=True)
                ref = WO.causal_wan_forward(sd, cfg, xc, t, text, cache=cache, cur_start_frame=3 * c, store_kv=store, **fwd)
                if not store:
                    errs.append(_rel(out, ref))
    torch.cuda.synchronize()
    print("wan 1.3B causal chunks 0..6, x0 rel-L2:", " ".join(f"{e:.2e}" for e in errs))
    assert max(errs) < 1e-2, errs
    # the caches agree too (layer 0 and the last layer, full length).  These are internal bf16 activations, not the x0 the
    # 1e-2 tolerance is defined on: after 29 bf16 blocks the keys / values of the last layer sit at ~1.2e-2 of the fp32
    # evaluation (measured 1.18e-2 / 1.16e-2; the final LayerNorm + head project most of that away: x0 is at 7e-3)
    kv = net._kv["pos"]
    for layer, bound in ((0, 5e-3), (29, 2e-2)):
        ek = _rel(kv["k"][layer, 0].float(), cache.k[layer][0].flatten(1))
        ev = _rel(kv["v"][layer, 0].float(), cache.v[layer][0].flatten(1))
        print(f"   KV cache layer {layer}: k {ek:.3e} v {ev:.3e}")
        assert ek < bound and ev < bound, (layer, ek, ev)


# ---------------------------------------------------------------------------------------------------------------------
# SDXL base configuration
# ---------------------------------------------------------------------------------------------------------------------
def test_sdxl_base_configuration_one_sample_eps_and_x0():
    from fastgen_b200.sdxl import StableDiffusionXL

    torch.manual_seed(0)
    with torch.device(DEV):
        net = StableDiffusionXL()
        for name, p in net.named_parameters():      # a non-degenerate random network: zero-initialised tensors refilled
            if p.dim() > 1 and float(p.detach().abs().max()) == 0.0:
                torch.nn.init.normal_(p, std=(p[0].numel()) ** -0.5)
        for p in net.parameters():                  # bf16-representable weights (the reference runs SDXL in bf16)
            p.data = p.data.bfloat16().float()
    net = net.eval()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    ocfg = dict(XO.SDXL_BASE)
    assert sum(v.numel() for k, v in sd.items() if k.startswith("unet.") and "logvar" not in k) == 2_567_463_684
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(1, 4, 128, 128, device=DEV, generator=g)
    text = torch.randn(1, 77, 2048, device=DEV, generator=g).bfloat16().float()
    pooled = torch.randn(1, 1280, device=DEV, generator=g).bfloat16().float()
    t = torch.tensor([0.749], dtype=torch.float64, device=DEV)
    with torch.no_grad(), OD.exact_fp32():
        for pred in ("eps", "x0"):
            out = net(x, t, condition=[text, pooled], fwd_pred_type=pred)
            ref = XO.sdxl_forward(sd, ocfg, x, t, text, pooled, fwd_pred_type=pred)
            # what the reference computes: the same network evaluated in bf16 (experiments/SDXL/config_dmd2.py: bf16)
            ref16 = XO.sdxl_forward(sd, ocfg, x.bfloat16(), t, text.bfloat16(), pooled.bfloat16(), fwd_pred_type=pred,
                                    dtype=torch.bfloat16)
            e, e16 = _rel(out, ref), _rel(ref16.float(), ref)
            print(f"sdxl base, 1 sample, {pred}: rel-L2 ours {e:.3e}, reference bf16 arithmetic {e16:.3e}")
            # 70 transformer blocks + 17 resnets on bf16 tensor-core operands: bounded by the tolerance or, where the
            # reference's own bf16 evaluation is further than that from the exact value, by its deviation
            assert e < max(1e-2, e16), (pred, e, e16)


def test_sd15_base_configuration_with_padding_masks():
    """SD 1.5 UNet at its published configuration (859.5 M parameters; heads of 40 / 80 / 160 channels), 2 samples at a
    64 x 64 latent with different prompt lengths"""
    from fastgen_b200.sdxl import StableDiffusion15

    torch.manual_seed(0)
    with torch.device(DEV):
        net = StableDiffusion15()
        for p in net.parameters():
            p.data = p.data.bfloat16().float()
    net = net.eval()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    ocfg = dict(XO.SD15_BASE)
    g = torch.Generator(device=DEV).manual_seed(6)
    x = torch.randn(2, 4, 64, 64, device=DEV, generator=g)
    text = torch.randn(2, 77, 768, device=DEV, generator=g).bfloat16().float()
    mask = torch.zeros(2, 77, device=DEV)
    mask[0, :77], mask[1, :11] = 1, 1
    t = torch.tensor([0.749, 0.3], dtype=torch.float64, device=DEV)
    with torch.no_grad(), OD.exact_fp32():
        out = net(x, t, condition=[text, mask], fwd_pred_type="eps")
        ref = XO.sd15_forward(sd, ocfg, x, t, text, mask, fwd_pred_type="eps")
        ref16 = XO.sd15_forward(sd, ocfg, x.bfloat16(), t, text.bfloat16(), mask.bfloat16(), fwd_pred_type="eps",
                                dtype=torch.bfloat16)
    e, e16 = _rel(out, ref), _rel(ref16.float(), ref)
    print(f"sd15 base, 2 samples with masks, eps: rel-L2 ours {e:.3e}, reference bf16 arithmetic {e16:.3e}")
    assert e < max(1e-2, e16), (e, e16)


# ---------------------------------------------------------------------------------------------------------------------
# EDM networks at full width, fp32 path (compute="bf16x3": three-product convolutions + fp32 attention)
# ---------------------------------------------------------------------------------------------------------------------
CIFAR = dict(img_resolution=32, img_channels=3, label_dim=10, sigma_shift=0.0, sigma_data=0.5, model_type="SongUNet",
             augment_dim=9, model_channels=128, channel_mult=[2, 2, 2], channel_mult_noise=1,
             embedding_type="positional", encoder_type="standard", decoder_type="standard", resample_filter=[1, 1],
             dropout=0.0, label_dropout=0, r_timestep=False, drop_precond=None)
IN64 = dict(img_resolution=64, img_channels=3, label_dim=1000, sigma_shift=0.0, sigma_data=0.5,
            model_type="DhariwalUNet", augment_dim=0, model_channels=192, channel_mult=[1, 2, 3, 4], channel_mult_emb=4,
            num_blocks=3, attn_resolutions=[32, 16, 8], dropout=0.0, label_dropout=0, r_timestep=False,
            drop_precond=None)


@pytest.mark.parametrize("name,cfg,batch", [("in64", IN64, 4), ("cifar", CIFAR, 8)])
def test_edm_full_width_fp32_path_within_1e_4(name, cfg, batch):
    from fastgen_b200.edm import EDMPrecond

    sd = {k: v.to(DEV) for k, v in EO.make_state_dict(cfg, 0).items()}
    net = EDMPrecond(**cfg, compute="bf16x3")
    net.load_state_dict(sd, strict=False)
    net = net.to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(5)
    r = cfg["img_resolution"]
    t = torch.tensor([80.0, 1.1, 0.3, 12.0, 0.05, 2.5, 40.0, 0.6][:batch], dtype=torch.float64, device=DEV)
    x = torch.randn(batch, 3, r, r, device=DEV, generator=g) * (t.float().view(-1, 1, 1, 1) ** 2 + 0.25).sqrt()
    labels = torch.eye(cfg["label_dim"], device=DEV)[torch.randint(cfg["label_dim"], (batch,), device=DEV, generator=g)]
    with torch.no_grad(), OD.exact_fp32():
        ref = EO.edm_precond_forward(sd, cfg, x, t, labels)
        out = net(x, t, condition=labels, fwd_pred_type="x0")
        raw_ref = EO.edm_precond_forward(sd, cfg, x, t, labels, return_raw=True)
    torch.cuda.synchronize()
    e = _rel(out, ref)
    per = ((out - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max().item()
    print(f"EDM {name} full width, bf16x3: x0 rel-L2 {e:.3e} (worst sample {per:.3e}), |F| {raw_ref.abs().mean():.3f}")
    assert e < 1e-4, e
    assert per < 2e-4, per
