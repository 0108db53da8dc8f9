"""Micro-benchmark of fg_peer_put + fg_peer_wait (the one-pass Ulysses exchange) on the Wan 1.3B shard shapes: fixed cost
(a one-row put) against payload-proportional cost, and against a plain peer-to-peer cudaMemcpy of the same bytes.
    torchrun --nproc-per-node P --master-addr 127.0.0.1 profiles/peer_micro.py      (gpurun --gpus P)"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastgen_b200 import _lib as L  # noqa: E402
from fastgen_b200 import ops  # noqa: E402
from fastgen_b200 import parallel as PL  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
enable = getattr(symm, "enable_symm_mem_for_group", None)
if enable is not None:
    enable(dist.group.WORLD.group_name)
D, P = 1536, world
Dl = D // P
stream = torch.cuda.current_stream().cuda_stream
lib = L.lib()


def bench(n, iters=200):
    """qkv exchange of a chunk of n tokens (n / P rows per rank): put + wait, `iters` times back to back"""
    n_l = n // P
    full_b = n * 3 * Dl * 2
    buf = symm.empty(full_b + 256, dtype=torch.uint8, device=dev)
    buf.zero_()
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    torch.cuda.synchronize()
    dist.barrier()
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    state = torch.zeros(4, dtype=torch.int64, device=dev)
    qkv = torch.randn(n_l, 3 * D, device=dev).bfloat16()
    p = ops.peer_put_params(qkv, PL.peer_put_qkv_desc(1, n, P, rank, D, Dl), ptrs, [q + full_b for q in ptrs], rank, state)

    def once():
        L.check(lib.fg_peer_put(C.byref(p), stream), "put")
        L.check(lib.fg_peer_wait(ptrs[rank] + full_b, P, state.data_ptr(), int(20e9), stream), "wait")

    for _ in range(5):
        once()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        once()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    sent = n_l * 3 * D * 2 * (P - 1) / P
    # plain P2P copy of the same number of bytes to the next rank, for the NVLink rate
    peer = hdl.get_buffer((rank + 1) % P, (full_b,), torch.uint8)
    src = torch.empty(int(sent), dtype=torch.uint8, device=dev)
    for _ in range(3):
        peer[:src.numel()].copy_(src)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(50):
        peer[:src.numel()].copy_(src)
    e1.record()
    torch.cuda.synchronize()
    us_cp = e0.elapsed_time(e1) / 50 * 1e3
    assert int(state[3].item()) == 0, "a wait timed out"
    if rank == 0:
        print(f"P={P} chunk of {n:6d} tokens ({n_l} rows/rank): put+wait {us:7.1f} us, {sent / 1e6:6.2f} MB to peers -> "
              f"{sent / us / 1e3:6.1f} GB/s;  plain P2P copy of the same bytes {us_cp:6.1f} us ({sent / us_cp / 1e3:6.1f} GB/s)",
              flush=True)
    dist.barrier()


for n in (P * 8, 4680, 4 * 4680, 32760 - 32760 % (P * 8)):
    bench(n)
dist.destroy_process_group()
