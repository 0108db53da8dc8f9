#!/usr/bin/env python
"""Benchmark of the few-step student sampler hot path (contract: see the task statement / DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): EDM ImageNet-64 class-conditional network (DhariwalUNet, 295.9 M parameters),
sCM-style 2-step sampler, t_list = [80.0, 1.1, 0.0], sde, batch 256 per GPU, random-init weights, synthetic inputs.
One "step" of this script = one full 2-step generation of one batch (2 network evaluations + the fused sampler
kernels).  N > 1: every rank generates its own batch of 256 (batch sharding of a larger sampling job, exactly how
the reference's FID script splits work; no collective on the data path) -> weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm's CPU port (the oracle) on the
host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IN64 = dict(img_resolution=64, img_channels=3, label_dim=1000, sigma_shift=0.0, sigma_data=0.5,
            model_type="DhariwalUNet", augment_dim=0, model_channels=192, channel_mult=[1, 2, 3, 4], channel_mult_emb=4,
            num_blocks=3, attn_resolutions=[32, 16, 8], dropout=0.0, label_dropout=0, r_timestep=False,
            drop_precond=None)
T_LIST = [80.0, 1.1, 0.0]
STEPS_PER_SAMPLE = 2
BATCH = 256
GFLOP_PER_SAMPLE_FORWARD = 219.33  # conv 206.84 + attention bmm 12.43 + linear 0.06
WORKLOAD = "EDM ImageNet-64 class-conditional sCM 2-step sampler, batch 256 per GPU (BASELINE configs[1])"
METRIC = "samples/sec, 2-step ImageNet-64 EDM student sampler"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["hbm_gbs"], d["bf16_tflops_sustained"], d["bf16_tflops"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def _randomize_zero_init_(net, seed: int) -> None:
    """The reference initialiser zeroes conv1 / proj / out_conv (EDM/network.py:607): refill them with
    N(0, 1/fan_in) so that the synthetic run exercises non-degenerate values (speed does not depend on it)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if p.dim() > 1 and float(p.abs().max()) == 0.0:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)


REF_DIR = os.environ.get("FASTGEN_B200_REF_DIR", os.path.join(ROOT, "baseline", "_ref"))


def _reference_package():
    """The UNMODIFIED reference installed by __graft_entry__.build() into baseline/_ref (git-ignored, travels to the GPU
    box), imported under oracle/ref_shims.py's inert stand-ins for its un-installable module-level imports.  None if the
    directory is absent (then the pinned CPU port under oracle/ is timed instead, and the line says so)."""
    if not os.path.isdir(os.path.join(REF_DIR, "fastgen")):
        return None
    try:
        os.environ["FASTGEN_REF"] = REF_DIR
        from oracle import ref_shims

        ref_shims.REF_ROOT = REF_DIR
        return ref_shims.load_reference()
    except Exception as e:  # pragma: no cover - reported in the JSON line
        print(f"[bench] reference import failed, timing the port: {e!r}", file=sys.stderr)
        return None


CPU_WORKLOADS = {
    # name: (network kwargs, steps, t_list, label_dim, resolution)
    "in64": (lambda: IN64, 2, T_LIST, 1000, 64),
    "cifar": (lambda: CIFAR, 1, None, 10, 32),
}


def cpu_reference_samples_per_sec(workload: str, sample_batch: int, threads: int, repeats: int = 1, warmup: int = 0):
    """The reference's own CPU path on a bounded sample of the workload: `fastgen.methods.model.FastGenModel.generator_fn`
    (methods/model.py:374-420) driving `fastgen.networks.EDM.network.EDMPrecond` in fp32 on `threads` host threads --
    unmodified reference code from baseline/_ref ("reference"); the oracle port ("port") only if that is absent.
    Returns (samples/s of the best repeat, mean seconds per repeat, kind)."""
    from oracle import edm_oracle as EO
    from oracle import sampler_oracle as SO

    torch.set_num_threads(threads)
    cfg_fn, steps, tl, ldim, res = CPU_WORKLOADS[workload]
    cfg = cfg_fn()
    sd = EO.make_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(sample_batch, 3, res, res, generator=g)
    labels = torch.eye(ldim)[torch.randint(ldim, (sample_batch,), generator=g)]
    ref = _reference_package()
    if ref is not None:
        from fastgen.methods.model import FastGenModel as RefModel
        from fastgen.networks.EDM.network import EDMPrecond as RefNet

        net = RefNet(**cfg)
        missing, unexpected = net.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        net = net.eval()

        def once():
            RefModel.generator_fn(net, noise, condition=labels, student_sample_steps=steps, t_list=tl,
                                  student_sample_type="sde")
        kind = "reference"
    else:
        eps = [torch.randn(sample_batch, 3, res, res, generator=g) for _ in range(steps - 1)]
        fwd = lambda x, t, **kw: EO.edm_precond_forward(sd, cfg, x, t, labels)  # noqa: E731

        def once():
            SO.generator_fn(fwd, SO.Schedule("edm"), noise, steps, tl, eps, "sde")
        kind = "port"
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            once()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sample_batch / min(times), sum(times) / len(times), kind


def _cpu_baseline(workload: str, sample: int, repeats: int = 1, warmup: int = 0):
    threads = os.cpu_count() or 1
    sps, mean_s, kind = cpu_reference_samples_per_sec(workload, sample, threads, repeats=repeats, warmup=warmup)
    steps = CPU_WORKLOADS[workload][1]
    what = ("the unmodified reference (baseline/_ref: FastGenModel.generator_fn + EDMPrecond, fp32)" if kind == "reference"
            else "the oracle (pinned CPU port of the reference; baseline/_ref absent)")
    return {"value": sps, "unit": "samples/s", "cores": threads, "kind": kind, "seconds_per_repeat": mean_s,
            "sample": f"{repeats} x one {steps}-step generation of a batch of {sample} (same network, t_list, fp32 "
                      f"weights) with {what}"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, same metric and
    config as our arm, each step a bounded sample (batch 16 of the 256; the whole CIFAR batch of 64)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = "cifar" if args.workload == "cifar" else "in64"
    sample = args.ref_sample or (64 if wl == "cifar" else 16)
    t0 = time.perf_counter()
    cpu = _cpu_baseline(wl, sample, repeats=max(1, args.steps), warmup=min(args.warmup, 1))
    metric = "samples/sec, 1-step CIFAR-10 EDM DMD2 generator" if wl == "cifar" else METRIC
    workload = ("EDM CIFAR-10 32x32 DMD2 one-step generator, batch 64 (BASELINE configs[0])" if wl == "cifar" else WORKLOAD)
    out = {
        "impl": "reference", "metric": metric, "value": cpu["value"], "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["seconds_per_repeat"] * 1000.0,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "t_list": T_LIST if wl == "in64" else "get_t_list(1)",
                   "sample": f"batch {sample} per step ({'the whole batch' if wl == 'cifar' else 'of 256'}), CPU fp32"},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(out))


class _Dist:
    """one process per GPU: rank / device / (optional) NCCL process group, the barrier and the max-over-ranks timer"""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])
        torch.cuda.synchronize()

    def timed(self, fn, iters):
        """K calls bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks"""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def close(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


def measure_in64(ctx: "_Dist", B: int, K: int, W: int, compute: str = "bf16", e2e: bool = True):
    """EDM ImageNet-64 2-step sampler at batch B per GPU: device-resident throughput, end-to-end throughput through
    FastGenModel.generator_fn with HOST buffers, and a per-launch CUDA-event pass for the roofline.  compute="bf16x3" is
    the fp32 path (three bf16 products per convolution + fp32 attention): the reference's precision default."""
    from fastgen_b200 import ops
    from fastgen_b200.edm import EDMPrecond
    from fastgen_b200.sampler import FastGenModel

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    torch.manual_seed(0)
    net = EDMPrecond(**IN64, compute=compute)
    _randomize_zero_init_(net, 0)
    net = net.to(dev).eval()
    torch.manual_seed(1 + rank)  # per-rank seeds, like set_random_seed(by_rank=True) in the reference scripts
    noise_h = torch.randn(B, 3, 64, 64).pin_memory()
    labels_h = torch.eye(1000)[torch.randint(1000, (B,))].pin_memory()
    noise, labels = noise_h.to(dev), labels_h.to(dev)
    out_h = torch.empty(B, 3, 64, 64).pin_memory()

    def gen(n_, l_):
        return FastGenModel.generator_fn(net, n_, condition=l_, student_sample_steps=STEPS_PER_SAMPLE, t_list=T_LIST,
                                         student_sample_type="sde")

    for _ in range(W):
        gen(noise, labels)
    ops.launch_count_reset()
    with ClockSampler(ctx.local) as clk:
        ms = ctx.timed(lambda: gen(noise, labels), K)
    res = {"value": world * B * K / (ms / 1000.0), "ms_per_step": ms / K, "gpu_launches": ops.launch_count(),
           "clocks": clk.summary(), "batch_per_gpu": B}

    if e2e:
        def e2e_step():
            n_d = noise_h.to(dev, non_blocking=True)
            l_d = labels_h.to(dev, non_blocking=True)
            out_h.copy_(gen(n_d, l_d), non_blocking=True)

        e2e_step()
        ms_e2e = ctx.timed(e2e_step, K)
        res["e2e"] = {"value": world * B * K / (ms_e2e / 1000.0), "unit": "samples/s", "ms_per_step": ms_e2e / K,
                      "h2d_bytes_per_step": noise_h.numel() * 4 + labels_h.numel() * 4,
                      "d2h_bytes_per_step": out_h.numel() * 4}

    # ---- roofline pass: CUDA-event pair around every launch of the same K steps ----
    hbm, tf_sus, tf_burst, peak_src = _peaks()
    net.profile = {}
    ctx.barrier()
    for _ in range(K):
        gen(noise, labels)
    ctx.barrier()
    prof, net.profile = net.profile, None
    tot_ms = sum(d["ms"] for d in prof.values())
    conv = prof.get("fg_conv_gemm", {"launches": 0, "ms": 0.0, "flops": 0.0})
    conv_tflops = conv["flops"] / (conv["ms"] / 1000.0) / 1e12 if conv["ms"] else 0.0
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path) and compute == "bf16" and B == BATCH:
        with open(tr_path) as f:
            traffic = json.load(f).get("fg_conv_gemm_bytes_per_launch")
    att = [d for k, d in prof.items() if k.startswith("fg_attention") and d.get("ms")]
    res["roofline"] = {
        "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv/linear)", "bound": "tensor",
        "achieved": conv_tflops, "peak": tf_sus, "unit": "TFLOP/s", "frac": conv_tflops / tf_sus if tf_sus else None,
        "traffic": traffic,
        "traffic_note": "DRAM read+write bytes per fg_conv_gemm launch (ncu, profiles/traffic.json); algorithmic bytes per "
                        "launch at batch 256 = 558 MB of activations (conv0 -> norm1 tensors in bf16; 610 MB with fp32 ones) + ~5 MB of weights (DESIGN.md): the excess is < 0 "
                        "because consumers hit the 126 MB L2",
        "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
        "how": "sum of algorithmic conv/linear FLOPs of all fg_conv_gemm launches / sum of their CUDA-event durations, "
               f"separate measurement pass of the same {K} steps"
               + (" (bf16x3: the tensor pipe executes 3x the algorithmic FLOPs, `executed_tflops`)" if compute == "bf16x3" else ""),
        "launches_per_step": conv["launches"] / K, "avg_launch_ms": conv["ms"] / max(conv["launches"], 1),
        "share_of_step": conv["ms"] / tot_ms if tot_ms else None,
        "per_kernel_ms_per_step": {k: round(d["ms"] / K, 3) for k, d in sorted(prof.items())},
        "attention_tflops": (sum(d["flops"] for d in att) / (sum(d["ms"] for d in att) / 1000.0) / 1e12) if att else None,
        "model_flops_utilisation": (world * B * K * STEPS_PER_SAMPLE * GFLOP_PER_SAMPLE_FORWARD * 1e9 /
                                    (ms / 1000.0) / 1e12) / (world * tf_sus),
    }
    if compute == "bf16x3":
        res["roofline"]["executed_tflops"] = 3.0 * conv_tflops
    del net
    torch.cuda.empty_cache()
    return res


class _Watchdog:
    """The secondary measurements of `also` must never cost the headline line: if they overrun (a hung collective, a
    box slower than planned) rank 0 prints the line it already has and every rank leaves with status 0."""

    def __init__(self, seconds: float, on_expire):
        self.timer = threading.Timer(seconds, self._fire)
        self.timer.daemon = True
        self.on_expire = on_expire

    def _fire(self):
        try:
            self.on_expire()
        finally:
            sys.stdout.flush()
            os._exit(0)

    def __enter__(self):
        self.timer.start()
        return self

    def __exit__(self, *exc):
        self.timer.cancel()


def run_ours(args):
    ctx = _Dist()
    world, rank = ctx.world, ctx.rank
    B, K, W = args.batch, args.steps, max(args.warmup, 0)
    main = measure_in64(ctx, B, K, W)

    # ---- CPU baseline (rank 0, N = 1 only): the unmodified reference on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = _cpu_baseline("in64", 16, repeats=2, warmup=0)

    line = {
        "metric": METRIC, "value": main["value"], "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "t_list": T_LIST,
                   "sample_type": "sde", "parallelism": f"batch-sharded x{world}, no data-path collective",
                   "l2": "no explicit flush: per step the kernels stream ~0.6 GB of weights and >10 GB of "
                         "activations, far beyond the 126 MB L2"},
        "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"], "roofline": main["roofline"],
        "cpu_baseline": cpu, "also": {},
    }
    also = line["also"]

    def emit():
        if rank == 0:
            print(json.dumps(line))
            sys.stdout.flush()

    if args.no_also:
        emit()
        ctx.close()
        return

    def expired():
        also["watchdog"] = f"secondary measurements cut off after {args.also_budget_s} s; entries present are complete"
        emit()

    with _Watchdog(args.also_budget_s, expired):
        # (1) the fp32 path (compute="bf16x3": what matches the reference at the reference's precision, <= 1e-4)
        try:
            r = measure_in64(ctx, B, max(2, K // 4), 1, compute="bf16x3", e2e=False)
            also["in64_bf16x3"] = {"metric": METRIC + " (fp32 path: three-product convolutions + fp32 attention, parity <= 1e-4)",
                                   "value": r["value"], "unit": "samples/s", "ms_per_step": r["ms_per_step"],
                                   "batch_per_gpu": B, "roofline": r["roofline"], "gpu_launches": r["gpu_launches"]}
        except Exception as e:  # pragma: no cover
            also["in64_bf16x3"] = {"error": repr(e)}
        # (2) strong scaling of BASELINE configs[1] as it is worded ("batch 256, batch-sharded over 8 GPUs": 256 / N per GPU)
        if world > 1 and BATCH % world == 0:
            try:
                r = measure_in64(ctx, BATCH // world, K, W, e2e=False)
                also["in64_strong_scaling"] = {"metric": METRIC + f", global batch {BATCH} sharded over {world} GPUs",
                                               "value": r["value"], "unit": "samples/s", "ms_per_step": r["ms_per_step"],
                                               "batch_per_gpu": BATCH // world, "scaling": "strong",
                                               "conv_tflops": r["roofline"]["achieved"],
                                               "model_flops_utilisation": r["roofline"]["model_flops_utilisation"]}
            except Exception as e:  # pragma: no cover
                also["in64_strong_scaling"] = {"error": repr(e)}
        keys = ("metric", "value", "unit", "ms_per_step", "scaling", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline")
        # (3) north_star's scaling target: Wan 2.1 1.3B 4-step self-forcing (configs[3]) under Ulysses sequence parallelism.
        # The metric is samples/s, so the N GPUs run N / 2 sequence-parallel groups of 2 (measured: 0.85 of linear at
        # every N; groups of 4 give the lower LATENCY per sample -- next entry -- but only 0.62 of linear, their 1 170-row
        # shards leave one GEMM tile per SM and pay the exchange's fixed cost on half as much work)
        sp_thr = 2 if world >= 2 else 1
        try:
            w = measure_wan(ctx, big=False, K=2, W=2, sp=sp_thr)
            if w is not None:
                also["wan13b"] = {k: w[k] for k in keys if k in w}
        except Exception as e:  # pragma: no cover
            also["wan13b"] = {"error": repr(e)}
        # (3a) lowest latency: one group of min(N, 4) ranks per sample
        if world >= 4:
            try:
                w = measure_wan(ctx, big=False, K=2, W=2, profile=False, e2e=False, sp=4)
                if w is not None:
                    also["wan13b_sp4_latency"] = {k: w[k] for k in keys if k in w}
            except Exception as e:  # pragma: no cover
                also["wan13b_sp4_latency"] = {"error": repr(e)}
        # (3b) two samples per sequence-parallel group (SURVEY 8d lists B = 1, 2, 8 for this config): larger shards per rank
        try:
            w = measure_wan(ctx, big=False, K=2, W=2, profile=False, e2e=False, batch=2, sp=sp_thr)
            if w is not None:
                also["wan13b_batch2"] = {k: w[k] for k in keys if k in w}
        except Exception as e:  # pragma: no cover
            also["wan13b_batch2"] = {"error": repr(e)}
        # (4) configs[0]: CIFAR-10 one-step generator, batch 64 -- one GPU by definition (rank 0; the others wait)
        try:
            if rank == 0:
                c = measure_cifar(64, 3, 3, cpu=(world == 1 and not args.no_cpu_baseline), device=ctx.local)
                also["cifar"] = {k: c[k] for k in keys if k in c}
            ctx.barrier()
        except Exception as e:  # pragma: no cover
            also["cifar"] = {"error": repr(e)}
        # (5) configs[2]: SDXL 4-step, batch 16 per GPU, batch-sharded (weak scaling like the headline)
        try:
            x = measure_sdxl(ctx, 16, 2, 1)
            if x is not None:
                also["sdxl"] = {k: x[k] for k in keys if k in x}
        except Exception as e:  # pragma: no cover
            also["sdxl"] = {"error": repr(e)}
        # (6) configs[4]: Wan 14B 720p 4-step, Ulysses over all N ranks (one sample timed after one warm-up sample)
        try:
            w = measure_wan(ctx, big=True, K=1, W=1, profile=False, e2e=False)
            if w is not None:
                also["wan14b"] = {k: w[k] for k in keys if k in w}
        except Exception as e:  # pragma: no cover
            also["wan14b"] = {"error": repr(e)}
    emit()
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------
# secondary workload (not the driver's headline line): Wan 2.1 1.3B T2V 480p 81-frame 4-step self-forcing
# (BASELINE configs[3]).  N GPUs = sequence-parallel groups of min(N, 4) ranks (12 heads) x N / min(N, 4) replicas.
# ---------------------------------------------------------------------------------------------------------------
WAN_TL = [0.999, 0.937, 0.833, 0.624, 0.0]
WAN_TFLOP_PER_SAMPLE = 995.0  # causal AR sample, 7 chunks x 5 passes (upper bound)


WAN14B_TFLOP_PER_SAMPLE = 4 * 6523.3  # bidirectional 14B forward at 720p (75 600 tokens) x 4 steps


def run_wan(args):
    ctx = _Dist()
    line = measure_wan(ctx, big=args.workload == "wan14b", K=args.steps, W=max(args.warmup, 0), batch=args.wan_batch,
                       sp=args.wan_sp)
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.close()


def measure_wan(ctx: "_Dist", big: bool, K: int, W: int, profile: bool = True, e2e: bool = True, batch: int = 1,
                sp: int = 0):
    """big: BASELINE configs[4] (bidirectional 14B, 720p, 4-step ode, Ulysses over all ranks); else configs[3] (1.3B
    self-forcing, Ulysses groups of min(N, 4) ranks x N / 4 replicas).  Returns the JSON line (rank 0) or None."""
    from fastgen_b200.sampler import CausVidModel, FastGenModel
    from fastgen_b200.wan import CausalWan, Wan

    dist = ctx.dist
    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    if sp <= 0:
        sp = world if big else min(world, 4)
    assert world % sp == 0, (world, sp)
    frames = 21
    lh, lw = (90, 160) if big else (60, 104)
    torch.manual_seed(0)
    with torch.device(dev):
        if big:
            with torch.no_grad():
                torch.set_default_dtype(torch.bfloat16)   # 14 B parameters: build them in bf16 straight away
                try:
                    net = Wan(model_id_or_local_path=Wan.MODEL_ID_14B, enable_logvar_linear=False)
                finally:
                    torch.set_default_dtype(torch.float32)
        else:
            net = CausalWan(total_num_frames=frames, chunk_size=3, use_wan_official_sinusoidal=True,
                            enable_logvar_linear=False)
        torch.nn.init.normal_(net.transformer.proj_out.weight, std=0.02)
    net = net.to(torch.bfloat16).eval()
    replica = rank // sp
    if sp > 1:
        groups = [dist.new_group(list(range(g * sp, (g + 1) * sp))) for g in range(world // sp)]
        net.sp_group = groups[replica]
    torch.manual_seed(1 + replica)  # ranks of one sequence-parallel group draw identical noise
    noise_h = torch.randn(batch, 16, frames, lh, lw).bfloat16().pin_memory()
    text_h = torch.randn(batch, 512, 4096).bfloat16().pin_memory()
    out_h = torch.empty(batch, 16, frames, lh, lw, dtype=torch.bfloat16).pin_memory()
    noise, text = noise_h.to(dev), text_h.to(dev)

    def gen(n_, c_):
        if big:
            return FastGenModel.generator_fn(net, n_, student_sample_steps=4, t_list=WAN_TL, condition=c_,
                                             student_sample_type="ode")
        return CausVidModel.generator_fn(net, n_, student_sample_steps=4, t_list=WAN_TL, condition=c_,
                                         student_sample_type="sde")

    barrier, timed = ctx.barrier, ctx.timed

    from fastgen_b200 import ops

    for _ in range(W):
        gen(noise, text)
    ops.launch_count_reset()
    net.graph_launches = 0
    with ClockSampler(local) as clk:
        ms = timed(lambda: gen(noise, text), K)
    launches = ops.launch_count() + net.graph_launches   # direct launches + launches replayed through CUDA graphs
    replicas = (world // sp) * batch     # samples in flight: data-parallel groups x samples per group
    value = replicas * K / (ms / 1000.0)

    def e2e_step():
        out_h.copy_(gen(noise_h.to(dev, non_blocking=True), text_h.to(dev, non_blocking=True)), non_blocking=True)

    if not big:
        e2e_step()   # (the 14B sample takes ~30 GPU-seconds: its plans are warm already, skip the extra untimed pass)
    ms_e2e = timed(e2e_step, K) if e2e else None
    hbm, tf_sus, tf_burst, peak_src = _peaks()
    prof = {}
    if profile:
        net.profile = {}
        barrier()
        gen(noise, text)
        barrier()
        prof, net.profile = net.profile, None
    tot_ms = sum(d["ms"] for d in prof.values())
    att = prof.get("fg_flash_attn:self_attn", {"launches": 0, "ms": 0.0, "flops": 0.0})
    att_tf = att["flops"] / (att["ms"] / 1000.0) / 1e12 if att["ms"] else 0.0
    gemm_ms = sum(d["ms"] for k, d in prof.items() if k.startswith("fg_conv_gemm"))
    gemm_fl = sum(d["flops"] for k, d in prof.items() if k.startswith("fg_conv_gemm"))
    line = None
    if rank == 0:
        line = {
            "metric": ("samples/sec, 4-step Wan 2.1 14B T2V 720p 81 frames" if big else
                       "samples/sec, 4-step self-forcing Wan 2.1 1.3B T2V 480p 81 frames"), "value": value,
            "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True,
            "scaling": "strong" if world == sp else f"strong within a group of {sp}, weak across the {world // sp} groups",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("Wan 2.1 14B T2V 720p 81-frame 4-step (ode), batch 1, Ulysses sequence parallel (BASELINE configs[4])"
                                    if big else
                                    f"Wan 2.1 1.3B T2V 480p 81-frame 4-step self-forcing, batch {batch} per group (BASELINE configs[3])"),
                       "latent": [batch, 16, frames, lh, lw], "t_list": WAN_TL, "chunk_size": None if big else 3,
                       "parallelism": f"Ulysses sequence parallel x{sp}, {world // sp} group(s) x batch {batch}",
                       "l2": ("inputs exceed L2: 28 GB of bf16 weights and 75 600-token activations streamed per forward" if big
                              else "inputs exceed L2: 2.8 GB of weights + up to 6 GB of KV cache streamed per sample")},
            "e2e": ({"value": replicas * K / (ms_e2e / 1000.0), "unit": "samples/s", "ms_per_step": ms_e2e / K,
                     "h2d_bytes_per_step": noise_h.numel() * 2 + text_h.numel() * 2, "d2h_bytes_per_step": out_h.numel() * 2}
                    if e2e else None),
            "gpu_launches": launches, "clocks": clk.summary(),
            "roofline": {"kernel": "attention2_kernel (tcgen05 flash attention, self-attention over the KV cache)",
                         "bound": "tensor", "achieved": att_tf, "peak": tf_sus, "unit": "TFLOP/s",
                         "frac": att_tf / tf_sus if tf_sus else None, "traffic": None,
                         "peak_source": peak_src + ", sustained bf16", "share_of_step": att["ms"] / tot_ms if tot_ms else None,
                         "gemm_tflops": gemm_fl / (gemm_ms / 1000.0) / 1e12 if gemm_ms else None,
                         "per_kernel_ms": {k: round(d["ms"], 2) for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]},
                         "model_flops_utilisation": (value * (WAN14B_TFLOP_PER_SAMPLE if big else WAN_TFLOP_PER_SAMPLE))
                                                    / (world * tf_sus)},
            "cpu_baseline": None,
        }
    del net
    torch.cuda.empty_cache()
    return line


SDXL_TFLOP_PER_SAMPLE_FORWARD = 6.75  # counted from our own layer table (sum of 2*M*N*K of every GEMM/conv + 4*N*Nk*d
                                      # of every attention at a 128x128 latent, 77 text tokens), see DESIGN.md


def run_sdxl(args):
    ctx = _Dist()
    line = measure_sdxl(ctx, args.batch if args.batch != BATCH else 16, args.steps, max(args.warmup, 0))
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.close()


def measure_sdxl(ctx: "_Dist", B: int, K: int, W: int):
    """BASELINE configs[2]: SDXL 1024x1024 (128x128 latent) DMD2 4-step UNet generation, batch 16 per GPU, bf16."""
    from fastgen_b200 import ops
    from fastgen_b200.sampler import FastGenModel
    from fastgen_b200.sdxl import StableDiffusionXL

    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    torch.manual_seed(0)
    with torch.device(dev):
        net = StableDiffusionXL()
    net = net.to(torch.bfloat16).eval()
    torch.manual_seed(1 + rank)
    noise_h = torch.randn(B, 4, 128, 128).bfloat16().pin_memory()
    text_h = torch.randn(B, 77, 2048).bfloat16().pin_memory()
    pooled_h = torch.randn(B, 1280).bfloat16().pin_memory()
    out_h = torch.empty(B, 4, 128, 128, dtype=torch.bfloat16).pin_memory()
    noise, text, pooled = noise_h.to(dev), text_h.to(dev), pooled_h.to(dev)

    def gen(n_, c_):
        return FastGenModel.generator_fn(net, n_, student_sample_steps=4, condition=c_, student_sample_type="sde")

    barrier, timed = ctx.barrier, ctx.timed
    for _ in range(max(W, 1)):
        out = gen(noise, [text, pooled])
    assert torch.isfinite(out.float()).all(), "non-finite samples"
    ops.launch_count_reset()
    with ClockSampler(local) as clk:
        ms = timed(lambda: gen(noise, [text, pooled]), K)
    launches = ops.launch_count()
    value = world * B * K / (ms / 1000.0)

    def e2e_step():
        c = [text_h.to(dev, non_blocking=True), pooled_h.to(dev, non_blocking=True)]
        out_h.copy_(gen(noise_h.to(dev, non_blocking=True), c), non_blocking=True)

    e2e_step()
    ms_e2e = timed(e2e_step, K)
    hbm, tf_sus, tf_burst, peak_src = _peaks()
    net.profile = {}
    barrier()
    gen(noise, [text, pooled])
    barrier()
    prof, net.profile = net.profile, None
    tot_ms = sum(d["ms"] for d in prof.values())
    gemm_ms = sum(d["ms"] for k, d in prof.items() if k.startswith("fg_conv_gemm"))
    gemm_fl = sum(d["flops"] for k, d in prof.items() if k.startswith("fg_conv_gemm"))
    gemm_n = sum(d["launches"] for k, d in prof.items() if k.startswith("fg_conv_gemm"))
    gemm_tf = gemm_fl / (gemm_ms / 1000.0) / 1e12 if gemm_ms else 0.0
    line = None
    if rank == 0:
        line = {
            "metric": "samples/sec, 4-step DMD2 SDXL 1024x1024 UNet generation", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"SDXL 1024x1024 DMD2 4-step UNet generation, batch {B} per GPU bf16 (BASELINE configs[2])",
                       "latent": [B, 4, 128, 128], "t_list": "linspace(0.999, 0, 5)", "sample_type": "sde",
                       "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": "inputs exceed L2: 5.1 GB of bf16 weights and >10 GB of activations streamed per forward"},
            "e2e": {"value": world * B * K / (ms_e2e / 1000.0), "unit": "samples/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": 2 * (noise_h.numel() + text_h.numel() + pooled_h.numel()),
                    "d2h_bytes_per_step": 2 * out_h.numel()},
            "gpu_launches": launches, "clocks": clk.summary(),
            "roofline": {"kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv / linear)", "bound": "tensor",
                         "achieved": gemm_tf, "peak": tf_sus, "unit": "TFLOP/s", "frac": gemm_tf / tf_sus if tf_sus else None,
                         "traffic": None, "peak_source": peak_src + ", sustained bf16",
                         "how": "sum of algorithmic FLOPs of all fg_conv_gemm launches / sum of their CUDA-event durations, "
                                "separate profiled pass of one 4-step generation",
                         "launches_per_step": gemm_n, "share_of_step": gemm_ms / tot_ms if tot_ms else None,
                         "per_kernel_ms": {k: round(d["ms"], 2) for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:24]},
                         "per_kernel_tflops": {k: round(d["flops"] / d["ms"] / 1e9, 1) for k, d in prof.items() if d["flops"] and d["ms"]},
                         "model_flops_utilisation": (value * 4 * SDXL_TFLOP_PER_SAMPLE_FORWARD) / (world * tf_sus)},
            "cpu_baseline": None,
        }
    del net
    torch.cuda.empty_cache()
    return line


CIFAR = dict(img_resolution=32, img_channels=3, label_dim=10, sigma_shift=0.0, sigma_data=0.5, model_type="SongUNet",
             augment_dim=9, model_channels=128, channel_mult=[2, 2, 2], channel_mult_noise=1,
             embedding_type="positional", encoder_type="standard", decoder_type="standard", resample_filter=[1, 1],
             dropout=0.0, label_dropout=0, r_timestep=False, drop_precond=None)
CIFAR_GFLOP_PER_SAMPLE_FORWARD = 42.38


def run_cifar(args):
    print(json.dumps(measure_cifar(args.batch if args.batch != BATCH else 64, args.steps, max(args.warmup, 0),
                                   graphs=not args.no_graphs, cpu=not args.no_cpu_baseline)))


def measure_cifar(B: int, K: int, W: int, graphs: bool = True, cpu: bool = True, device: int = 0):
    """BASELINE configs[0]: EDM CIFAR-10 32x32 DMD2 one-step generator, batch 64 (the reference's CPU-runnable case)."""
    from fastgen_b200 import ops
    from fastgen_b200.edm import EDMPrecond
    from fastgen_b200.sampler import FastGenModel

    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    torch.manual_seed(0)
    net = EDMPrecond(**CIFAR)
    _randomize_zero_init_(net, 0)
    net = net.to(dev).eval()
    net.use_cuda_graphs = graphs   # 174 launches of 10-30 us: replay them as one CUDA graph
    torch.manual_seed(1)
    noise_h = torch.randn(B, 3, 32, 32).pin_memory()
    labels_h = torch.eye(10)[torch.randint(10, (B,))].pin_memory()
    out_h = torch.empty(B, 3, 32, 32).pin_memory()
    noise, labels = noise_h.to(dev), labels_h.to(dev)

    def gen(n_, l_):
        return FastGenModel.generator_fn(net, n_, condition=l_, student_sample_steps=1)

    def timed(fn, iters):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(W):
        gen(noise, labels)
    iters = K * 20     # a step is ~3 ms: time 20 generations per "step" so that the events see a long region
    ops.launch_count_reset()
    net.graph_launches = 0
    with ClockSampler(device) as clk:
        ms = timed(lambda: gen(noise, labels), iters)
    launches = ops.launch_count() + net.graph_launches   # direct launches + launches replayed through the CUDA graph

    def e2e_step():
        out_h.copy_(gen(noise_h.to(dev, non_blocking=True), labels_h.to(dev, non_blocking=True)), non_blocking=True)

    e2e_step()
    ms_e2e = timed(e2e_step, iters)
    hbm, tf_sus, tf_burst, peak_src = _peaks()
    value = B * iters / (ms / 1000.0)
    # roofline pass: every launch of 20 generations bracketed by CUDA events (eager launches: a graph has no per-node events)
    graphs, net.use_cuda_graphs = net.use_cuda_graphs, False
    gen(noise, labels)
    net.profile = {}
    torch.cuda.synchronize()
    for _ in range(20):
        gen(noise, labels)
    torch.cuda.synchronize()
    prof, net.profile = net.profile, None
    net.use_cuda_graphs = graphs
    tot_ms = sum(d["ms"] for d in prof.values())
    conv = prof.get("fg_conv_gemm", {"launches": 0, "ms": 0.0, "flops": 0.0})
    conv_tf = conv["flops"] / (conv["ms"] / 1000.0) / 1e12 if conv["ms"] else None
    cpu_base = _cpu_baseline("cifar", B, repeats=2, warmup=0) if cpu else None
    line = {
        "metric": "samples/sec, 1-step CIFAR-10 EDM DMD2 generator", "value": value, "unit": "samples/s", "n_gpus": 1,
        "steps": K, "warmup": W, "ms_per_step": ms / iters, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"EDM CIFAR-10 32x32 DMD2 one-step generator, batch {B} (BASELINE configs[0])",
                   "generations_timed": iters, "t_list": "get_t_list(1) = [79.56, 0]", "cuda_graph": bool(net.use_cuda_graphs),
                   "l2": "the whole working set (223 MB of fp32 weights as bf16 = 111 MB, ~100 MB of activations) is of the "
                         "order of the 126 MB L2: this configuration is launch-latency bound, not bandwidth bound"},
        "e2e": {"value": B * iters / (ms_e2e / 1000.0), "unit": "samples/s", "ms_per_step": ms_e2e / iters,
                "h2d_bytes_per_step": 4 * (noise_h.numel() + labels_h.numel()), "d2h_bytes_per_step": 4 * out_h.numel()},
        "gpu_launches": launches, "launches_per_generation": launches / iters, "clocks": clk.summary(),
        "roofline": {"kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv/linear)", "bound": "tensor", "achieved": conv_tf,
                     "peak": tf_burst, "unit": "TFLOP/s", "frac": conv_tf / tf_burst if conv_tf and tf_burst else None,
                     "traffic": None, "peak_source": peak_src + ", burst bf16 (30 us kernels timed one by one)",
                     "how": "sum of algorithmic conv/linear FLOPs of the fg_conv_gemm launches of 20 generations / sum of "
                            "their CUDA-event durations (eager launches; the timed region replays them as one CUDA graph)",
                     "launches_per_generation": conv["launches"] / 20, "avg_launch_ms": conv["ms"] / max(conv["launches"], 1),
                     "share_of_step": conv["ms"] / tot_ms if tot_ms else None,
                     "per_kernel_ms_per_generation": {k: round(d["ms"] / 20, 4) for k, d in sorted(prof.items())},
                     "model_flops_utilisation": value * CIFAR_GFLOP_PER_SAMPLE_FORWARD * 1e9 / 1e12 / tf_sus},
        "cpu_baseline": cpu_base,
    }
    del net
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="in64", choices=["in64", "cifar", "wan", "wan14b", "sdxl"],
                    help="in64 = the headline line the driver records; wan = secondary workload for profiles/")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample", type=int, default=0, help="--impl reference: batch of the bounded CPU sample (default 16 / 64)")
    ap.add_argument("--wan-batch", type=int, default=1, help="wan workload: samples per sequence-parallel group")
    ap.add_argument("--wan-sp", type=int, default=0, help="wan workload: ranks per sequence-parallel group (0: min(N, 4); 14B: N)")
    ap.add_argument("--no-also", action="store_true", help="in64: skip the secondary measurements (fp32 path, strong scaling, Wan 1.3B)")
    ap.add_argument("--also-budget-s", type=float, default=600.0, help="in64: wall-clock cap of the secondary measurements")
    ap.add_argument("--no-graphs", action="store_true", help="cifar workload: launch every kernel instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("wan", "wan14b"):
        run_wan(args)
    elif args.workload == "sdxl":
        run_sdxl(args)
    elif args.workload == "cifar":
        run_cifar(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
