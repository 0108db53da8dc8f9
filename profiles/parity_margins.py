"""Print the measured relative L2 errors of the network parity tests (the tests only assert a bound): how much margin
the 1e-2 bf16 tolerance has on this box.  usage (GPU box): python profiles/parity_margins.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm()).item()


def sdxl():
    import test_sdxl_parity_gpu as T
    from oracle import sdxl_oracle as XO

    for name, cfg, S in (("3level", XO.tiny_config(), 32),
                         ("2level_d64x3", XO.tiny_config(block_out_channels=(64, 192), down_attn=(False, True),
                                                         transformer_layers=(1, 2), num_heads=(1, 3), layers_per_block=1), 16)):
        net, sd = T._net(cfg, 0)
        x, text, pooled = T._inputs(cfg, 2, S, 1)
        t = torch.tensor([0.999, 0.4], dtype=torch.float64)
        for pred in ("eps", "x0"):
            vals = []
            for _ in range(3):
                out = net(x.cuda(), t.cuda(), condition=[text.cuda(), pooled.cuda()], fwd_pred_type=pred)
                vals.append(rel(out, XO.sdxl_forward(sd, cfg, x, t, text, pooled, fwd_pred_type=pred)))
            print(f"sdxl {name:14s} {pred:4s} rel-L2 {min(vals):.3e} .. {max(vals):.3e}   (bound 1e-2)")


def wan():
    import test_wan_parity_gpu as T
    from oracle import wan_oracle as WO

    cfg = WO.tiny_config()
    net, sd = T._nets(cfg, 0, "Wan")
    x, text = T._inputs(cfg, 2, 3, 8, 12)
    t = torch.tensor([0.8, 0.3], dtype=torch.float64)
    for pred in ("x0", "flow"):
        out = net(x.cuda(), t.cuda(), condition=text.cuda(), fwd_pred_type=pred)
        ref32 = WO.wan_forward(sd, cfg, x, t, text, fwd_pred_type=pred, dtype=torch.float32)
        print(f"wan bidirectional   {pred:4s} rel-L2 {rel(out, ref32):.3e}   (bound 1e-2)")


def dit():
    import test_dit_parity_gpu as T
    from oracle import dit_oracle as DO

    for cfg in (DO.tiny_config(), DO.tiny_config(hidden_size=144, num_heads=2, depth=3)):
        net, sd = T._net(cfg, 0)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(3, cfg["in_channels"], cfg["input_size"], cfg["input_size"], generator=g)
        t = torch.tensor([0.9, 0.5, 0.1], dtype=torch.float64)
        lab = torch.eye(cfg["num_classes"])[torch.tensor([1, 7, 3])]
        out = net(x.cuda(), t.cuda(), condition=lab.cuda(), fwd_pred_type="x0")
        print(f"dit d={cfg['hidden_size'] // cfg['num_heads']:3d}          x0   rel-L2 "
              f"{rel(out, DO.dit_forward(sd, cfg, x, t, lab, fwd_pred_type='x0')):.3e}   (bound 1e-2)")


if __name__ == "__main__":
    for fn in (sdxl, wan, dit):
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            print(fn.__name__, "FAILED:", repr(e))
