"""CPU tests of host-side logic that needs no kernel: checkpoint key conversions, schedules on the host, argument
validation of the network wrappers (the compute path itself raises without CUDA)."""
import pytest
import torch

from oracle import dit_oracle as DO
from oracle import wan_oracle as WO


def test_dit_facebook_checkpoint_key_conversion():
    """reference DiT/network.py:346-401: Facebook-DiT names map onto ours; pos_embed goes straight into the buffer"""
    from fastgen_b200.dit import DiT

    cfg = DO.tiny_config()
    sd = DO.make_state_dict(cfg, 2)
    fb = {}
    for k, v in sd.items():
        k = k.replace("t_embedder.proj_net", "t_embedder.mlp").replace(".attention.", ".attn.")
        k = k.replace(".feed_forward.", ".mlp.").replace("y_embedder.class_embeddings", "y_embedder.embedding_table")
        k = k.replace("final_layer.projection", "final_layer.linear")
        k = k.replace("final_layer.adaptive_params", "final_layer.adaLN_modulation")
        k = k.replace(".conditioning_net.", ".adaLN_modulation.")
        fb[k] = v
    fb["pos_embed"] = sd["pos_embed"] + 1.0
    net = DiT(**cfg)
    missing, unexpected = net.load_state_dict(fb, strict=False)
    assert missing == ["pos_embed"] and not unexpected
    for k, v in sd.items():
        want = v + 1.0 if k == "pos_embed" else v
        assert torch.equal(net.state_dict()[k], want), k
    assert set(net.state_dict().keys()) == set(sd.keys())


def test_dit_positional_embedding_matches_oracle_restatement():
    from fastgen_b200.dit import DiT

    net = DiT(**DO.tiny_config())
    assert torch.equal(net.pos_embed, DO.pos_embed_2d(128, 4))


def test_networks_refuse_cpu_tensors_and_training_outputs():
    from fastgen_b200 import _lib
    from fastgen_b200.dit import DiT
    from fastgen_b200.wan import CausalWan, Wan

    dit = DiT(**DO.tiny_config())
    with pytest.raises(_lib.FastGenB200Error):
        dit(torch.zeros(1, 4, 8, 8), torch.full((1,), 0.5), condition=torch.zeros(1, 10))
    with pytest.raises(NotImplementedError):
        dit(torch.zeros(1, 4, 8, 8), torch.full((1,), 0.5), condition=torch.zeros(1, 10), return_logvar=True)
    cfg = WO.tiny_config(layers=1)
    wan = Wan(transformer_config=cfg)
    with pytest.raises(_lib.FastGenB200Error):
        wan(torch.zeros(1, 16, 2, 8, 8), torch.full((1,), 0.5), condition=torch.zeros(1, 8, 64))
    assert wan(torch.zeros(1, 16, 2, 8, 8), torch.full((1,), 0.5), return_features_early=True) == []
    with pytest.raises(NotImplementedError):
        wan(torch.zeros(1, 16, 2, 8, 8), torch.full((1,), 0.5), condition=torch.zeros(1, 8, 64), feature_indices={0})
    with pytest.raises(AssertionError):
        wan(torch.zeros(1, 16, 2, 8, 8), torch.full((1,), 0.5), condition=torch.zeros(1, 8, 64), fwd_pred_type="nope")
    with pytest.raises(ValueError):
        wan(torch.zeros(1, 16, 2, 8, 8), torch.full((1,), 0.5), condition=torch.zeros(1, 8, 64), r=torch.ones(1))
    wr = Wan(transformer_config=cfg, r_timestep=True, r_embedder_init="random")      # reference Wan.init_embedder
    keys = set(wr.state_dict().keys())
    assert {"transformer.r_embedder.time_embedder.linear_1.weight", "transformer.r_embedder.time_proj.bias"} <= keys
    assert not any("r_embedder.text_embedder" in k for k in keys)
    assert float(wr.transformer.r_embedder.time_proj.weight.abs().max()) == 0.0        # zero r projection initially
    wp = Wan(transformer_config=cfg, r_timestep=True)                                  # "pretrained": a copy
    assert torch.equal(wp.transformer.r_embedder.time_proj.weight, wp.transformer.condition_embedder.time_proj.weight)
    with pytest.raises(ValueError):
        Wan(transformer_config=cfg, r_timestep=True, time_cond_type="bogus")
    cw = CausalWan(transformer_config=cfg, chunk_size=2, total_num_frames=5)
    assert cw.chunk_size == 2 and cw.total_num_frames == 5 and cw.net_pred_type == "flow" and cw.schedule_type == "rf"
    cw.clear_caches()  # no caches yet: a no-op
    with pytest.raises(ValueError):
        Wan(model_id_or_local_path="Wan-AI/unknown")
    assert Wan.get_model_id("Wan-AI/Wan2.1-T2V-14B-Diffusers") == Wan.MODEL_ID_14B


def test_wan_rope_pair_tables_match_oracle_tables():
    from fastgen_b200.wan import _rope_pair_tables

    cfg = WO.tiny_config()
    cos, sin, nt, nh = _rope_pair_tables(cfg["head_dim"], cfg["rope_max_seq_len"])
    fc, fs = WO.rope_tables(cfg)
    assert torch.equal(cos, fc[:, 0::2]) and torch.equal(sin, fs[:, 1::2])
    assert (nt, nh) == (22, 21)


def test_sinusoid_variants_match_oracle():
    from fastgen_b200.wan import _sinusoid

    t = torch.tensor([999.0, 937.0, 0.0], dtype=torch.float64)
    assert torch.equal(_sinusoid(t, 256, True), WO.sinusoidal_embedding_1d_wan(256, t))
    assert torch.equal(_sinusoid(t, 256, False), WO.timesteps_diffusers(t, 256))


def test_trig_schedule_scalars_match_reference_fixture():
    """TrigNoiseSchedule host logic against values of the unmodified reference (oracle/make_golden.py:trig_fixture)"""
    import os

    import numpy as np

    from fastgen_b200.noise_schedule import get_noise_schedule

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "schedule_trig.npz"))
    ns = get_noise_schedule("trig")
    t = torch.from_numpy(g["t"])
    assert torch.equal(ns.alpha(t), torch.from_numpy(g["alpha"])) and torch.equal(ns.sigma(t), torch.from_numpy(g["sigma"]))
    assert torch.equal(ns.sqrt_snr(t), torch.from_numpy(g["sqrt_snr"]))
    assert torch.equal(ns.get_t_list(4), torch.from_numpy(g["t_list4"])) and ns.max_sigma == float(g["max_sigma"])
    assert torch.equal(ns.sigma_idx_to_t(torch.tensor([0, 17, 999])), torch.from_numpy(g["idx_t"]))
    assert torch.equal(ns.sqrt_snr_to_t(torch.tensor([0.1, 1.0, 30.0], dtype=torch.float64)), torch.from_numpy(g["snr_t"]))
    assert torch.equal(ns.rescale_t(t), t)
    with pytest.raises(AssertionError):
        ns.alpha(torch.tensor([1.7], dtype=torch.float64))


def test_package_root_exports_resolve():
    import fastgen_b200 as F

    for name in F.__all__:
        assert getattr(F, name) is not None
    with pytest.raises(AttributeError):
        F.no_such_thing


def test_original_wan_checkpoint_key_remap_matches_the_reference(monkeypatch):
    """`Wan.load_state_dict` on original-format checkpoints (flat / nested under 'generator', 'state_dict', 'model'; keys
    with 'model.' / 'module.' prefixes) hands nn.Module.load_state_dict exactly the keys the UNMODIFIED reference does
    (tests/golden/wan_original_key_remap.json, oracle/make_golden.py:wan_remap_fixture); every remapped key exists in the
    network, and diffusers-named checkpoints pass through untouched."""
    import json
    import os

    import torch

    import fastgen_b200.wan as W
    from oracle import wan_oracle as WO

    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wan_original_key_remap.json")))
    net = W.Wan(transformer_config=WO.tiny_config(layers=2), use_wan_official_sinusoidal=True, enable_logvar_linear=False)
    own = set(net.state_dict().keys())
    got = {}
    monkeypatch.setattr(torch.nn.Module, "load_state_dict", lambda self, sd, **kw: got.update(keys=list(sd.keys()), kw=sorted(kw)))
    for tag, rec in fx.items():
        src = {k: torch.zeros(1) for k in rec["src"]}
        kw = {"strict": False} if rec["kw"] else {}
        net.load_state_dict({rec["nest"]: src} if rec["nest"] else src, **kw)
        assert got["keys"] == rec["dst"], tag
        assert got["kw"] == rec["kw"], tag
        assert set(got["keys"]) <= own, (tag, sorted(set(got["keys"]) - own)[:5])
    assert len(fx["flat"]["dst"]) == len(own)      # the original T2V key list covers every parameter of the network


def test_dit_facebook_key_conversion_matches_the_reference_record():
    """same tagged Facebook-DiT checkpoint through our `DiT.load_state_dict`: every tensor lands under the key the
    UNMODIFIED reference put it (tests/golden/dit_facebook_key_remap.json, oracle/make_golden.py:dit_remap_fixture), the
    returned missing / unexpected lists agree, a pos_embed of the wrong shape is ignored like the reference does"""
    import json
    import os

    from fastgen_b200.dit import DiT

    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dit_facebook_key_remap.json")))
    net = DiT(**DO.tiny_config())
    own = net.state_dict()
    assert set(own.keys()) == set(fx["tag_of"].keys())
    from oracle.make_golden import facebook_dit_keys

    fb_shapes = facebook_dit_keys(own)
    assert list(fb_shapes.keys()) == fx["src"]
    tagged = {k: torch.full_like(fb_shapes[k], float(i + 1)) for i, k in enumerate(fx["src"])}
    res = net.load_state_dict(tagged, strict=False)
    assert list(res.missing_keys) == fx["missing"] and list(res.unexpected_keys) == fx["unexpected"]
    after = net.state_dict()
    assert {k: float(v.flatten()[0]) for k, v in after.items()} == fx["tag_of"]
    bad = dict(tagged)
    bad["pos_embed"] = torch.full((1, 3, 5), 99.0)
    res2 = net.load_state_dict(bad, strict=False)
    assert list(res2.missing_keys) == fx["bad_pos_missing"] and float(net.pos_embed.flatten()[0]) == fx["bad_pos_tag"]
