import os, sys, torch
sys.path.insert(0, "/root/repo")
from bench import IN64, _randomize_zero_init_
from fastgen_b200.edm import EDMPrecond
B=256
x = torch.randn(B, 3, 64, 64, device="cuda") * 80
t = torch.full((B,), 80.0, device="cuda", dtype=torch.float64)
lab = torch.eye(1000, device="cuda")[torch.randint(1000, (B,), device="cuda")]
for fuse in (True, False, True, False):
    net = EDMPrecond(**IN64)
    _randomize_zero_init_(net, 0)
    net.fuse_gn_stats = fuse
    net = net.cuda().eval()
    for _ in range(3): net(x, t, condition=lab)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net(x, t, condition=lab)
    e1.record(); torch.cuda.synchronize()
    print("fuse_gn_stats", fuse, "forward ms", e0.elapsed_time(e1)/10, flush=True)
    del net
