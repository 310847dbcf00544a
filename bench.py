#!/usr/bin/env python
"""Benchmark of the hot path: one full G+D training iteration of the MNIST conv CounteRGAN
(conditional_counteRGAN/mnist/trainer.py:89-132) at batch 512 per GPU, synthetic 1x28x28 inputs.

    python bench.py --gpus N --steps K --warmup W            # native arm (libpcg, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the
                                                             # reference's step on the host cores

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM, `e2e` = the same
through the user-facing trainer call with pinned host batches copied in and the loss scalars read
back every step.  See DESIGN.md §Measurement for how every field is produced.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "conditional_counteRGAN/mnist conv CounteRGAN, frozen CNN classifier, synthetic 1x28x28, batch 512/GPU"
METRIC = "GAN train samples/sec (G+D step) on MNIST CounteRGAN"
BATCH = 512
CPU_SAMPLE_BATCH = 128          # reference default batch_size (config.py:4); bounds the CPU arm's step time

# algorithmic work of the dominant kernel: 64->64 3x3 conv over B*784 pixels
CONV_FLOPS = 2.0 * BATCH * 784 * 64 * 576


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_step_rate(steps, warmup, batch=CPU_SAMPLE_BATCH, threads=None):
    """Times the oracle port of the reference step (oracle/mnist_countergan.py) on the host cores.  Fallback of
    ``reference_step_rate`` when ``baseline/_ref`` was not staged."""
    import torch
    from oracle import mnist_countergan as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    PG = O.synth_params(O.g_param_shapes(), 1, "G")
    PD = O.synth_params(O.d_param_shapes(), 2, "D")
    PC = O.synth_params(O.c_param_shapes(), 3, "C")
    S = O.make_state(PG, O.g_buffers(), PD, PC)
    batches = [O.synth_batch(batch, 10 + i) for i in range(2)]
    for i in range(warmup):
        O.countergan_step(S, *batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        O.countergan_step(S, *batches[i % 2])
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, threads


class _TimedLoader:
    """The ``train_loader`` handed to the reference's train_countergan: yields ``warmup + steps`` synthetic (x, y)
    batches and stamps the wall clock each time the loop asks for the next one, i.e. exactly between two iterations of
    trainer.py:89.  Stops early (never before ``min_steps`` timed iterations) once ``budget_s`` is used up."""

    def __init__(self, batches, warmup, steps, budget_s, min_steps=3):
        self.batches, self.warmup, self.steps, self.budget_s, self.min_steps = batches, warmup, steps, budget_s, min_steps
        self.stamps = []
        self.sync = None                          # device synchronisation before each stamp (GPU comparator runs)

    def __iter__(self):
        t_begin = time.perf_counter()
        for i in range(self.warmup + self.steps):
            if self.sync is not None:
                self.sync()
            now = time.perf_counter()
            self.stamps.append(now)
            if i >= self.warmup + self.min_steps and now - t_begin > self.budget_s:
                return
            yield self.batches[i % len(self.batches)]
        if self.sync is not None:
            self.sync()
        self.stamps.append(time.perf_counter())

    def timed(self):
        """(iterations, seconds) of the timed region: from the request of batch ``warmup`` to the last stamp."""
        n = len(self.stamps) - 1 - self.warmup
        return n, self.stamps[-1] - self.stamps[self.warmup]


def reference_step_rate(steps, warmup, batch=BATCH, threads=None, budget_s=240.0, device="cpu", autocast=None):
    """Drives the UNMODIFIED reference ``train_countergan`` (conditional_counteRGAN/mnist/trainer.py:76-163, staged at
    baseline/_ref by baseline/vendor_ref.py) with the reference's own modules on the host cores: device="cpu", fp32,
    hyper-parameters of config.py:10-15, synthetic batches of the bench workload.  Returns None when nothing is staged.
    matplotlib (trainer.py:5) is absent from this image and stubbed; nothing else is touched."""
    import contextlib
    import importlib
    import tempfile
    import types
    from unittest import mock
    import torch
    from baseline import vendor_ref
    from oracle import mnist_countergan as O      # synthetic batch generator only
    d = vendor_ref.staged_dir()
    if d is None:
        return None
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, mock.MagicMock())
    for k in [k for k in sys.modules if k in ("trainer", "models", "config") or k.startswith("models.")]:
        del sys.modules[k]
    # models/discriminator.py:3,6 reads ONE default from the experiment's config module (Config.num_classes); config.py
    # itself is not staged (it holds a hard-coded API key), so the bench supplies the values of config.py:4-27 it needs
    cfg_mod = types.ModuleType("config")
    cfg_mod.Config = type("Config", (), dict(batch_size=128, d_lr=1e-5, g_lr=5e-5, lambda_adv=1.0, lambda_cls=1.0,
                                             lambda_reg=2.5, lambda_mask=2.0, patch_size=7, num_modifiable_patches=10,
                                             img_shape=(1, 28, 28), num_classes=10))
    sys.modules["config"] = cfg_mod
    sys.path.insert(0, d)
    try:
        trainer = importlib.import_module("trainer")
        Gm = importlib.import_module("models.generator")
        Dm = importlib.import_module("models.discriminator")
        Cm = importlib.import_module("models.classifier")
    finally:
        sys.path.remove(d)
    torch.manual_seed(0)
    G, D, C = Gm.ResidualGenerator(), Dm.Discriminator(), Cm.CNNClassifier().eval()      # main.py:24-33
    for p in C.parameters():
        p.requires_grad = False
    batches = [O.synth_batch(batch, 10 + i)[:2] for i in range(2)]
    loader = _TimedLoader(batches, warmup, steps, budget_s)
    with tempfile.TemporaryDirectory() as tmp:
        K = cfg_mod.Config
        cfg = types.SimpleNamespace(g_lr=K.g_lr, d_lr=K.d_lr, num_epochs_gan=1, num_classes=K.num_classes,
                                    patch_size=K.patch_size, num_modifiable_patches=K.num_modifiable_patches,
                                    lambda_adv=K.lambda_adv, lambda_cls=K.lambda_cls, lambda_reg=K.lambda_reg,
                                    lambda_mask=K.lambda_mask, save_dir=tmp, generator_path=os.path.join(tmp, "generator.pt"))
        if device != "cpu":                                   # informational comparator: torch eager on the GPU
            G, D, C = G.to(device), D.to(device), C.to(device)
            loader.sync = torch.cuda.synchronize
        ctx = torch.autocast("cuda", dtype=autocast) if autocast is not None else contextlib.nullcontext()
        with contextlib.redirect_stdout(sys.stderr), ctx:     # the reference prints; stdout carries the JSON line only
            trainer.train_countergan(G, D, C, loader, cfg, device)
    n, dt = loader.timed()
    return batch * n / dt, dt / n * 1e3, threads, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    ref = reference_step_rate(steps, warmup)
    if ref is not None:
        rate, ms, threads, done = ref
        kind, batch = "reference", BATCH
        sample = (f"unmodified reference train_countergan (trainer.py:76-163, torch CPU fp32, staged at baseline/_ref), "
                  f"{done} timed iterations of B={BATCH} after {warmup} warm-up, {threads} threads")
    else:                                          # baseline/_ref not staged: the oracle port, bounded
        done, warmup = min(steps, 40), min(warmup, 3)
        rate, ms, threads = cpu_step_rate(done, warmup)
        kind, batch = "port", CPU_SAMPLE_BATCH
        sample = (f"oracle port of trainer.py:89-132 (torch CPU fp32), {done} steps of B={CPU_SAMPLE_BATCH} "
                  f"after {warmup} warm-up, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": done, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": batch, "cpu_batch_per_step": batch},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if done != steps:
        line["steps_requested"] = steps           # the CPU arm stops after ~4 minutes of timed iterations
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    import pcg_b200  # noqa: F401
    from pcg_b200 import _lib
    from pcg_b200.mnist import trainer as T
    from pcg_b200.mnist.models.classifier import CNNClassifier
    from pcg_b200.mnist.models.discriminator import Discriminator
    from pcg_b200.mnist.models.generator import ResidualGenerator
    from oracle import mnist_countergan as O      # synthetic batch generator only (inputs, not compute)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: libpcg has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()

    torch.manual_seed(0)                          # identical replicas on every rank
    G = ResidualGenerator().to(dev)
    D = Discriminator().to(dev)
    C = CNNClassifier().to(dev).eval()
    for p in C.parameters():
        p.requires_grad = False
    import types
    cfg = types.SimpleNamespace(g_lr=5e-5, d_lr=1e-5, num_classes=10, patch_size=7, num_modifiable_patches=10,
                                lambda_adv=1.0, lambda_cls=1.0, lambda_reg=2.5, lambda_mask=2.0)
    tr = T.CounterGanTrainer(G, D, C, cfg, dev, precision=args.precision, use_graph=not args.no_graph)
    T._warm_plan(tr, BATCH)

    # ring of device-resident synthetic batches (per-rank seeds), cycled so no step reuses the previous inputs
    ring = []
    for i in range(4):
        x, y, t, m = O.synth_batch(BATCH, 1000 * (rank + 1) + i, mnist_like=(i % 2 == 1))
        ring.append(tuple(v.to(dev).contiguous() for v in (x, y, t, m)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launches per step (counted on an eager, un-captured step)
    tr_eager_graph = tr.use_graph
    tr.use_graph = False
    n0 = _lib.launch_count()
    tr.step(*ring[0])
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - n0
    tr.use_graph = tr_eager_graph

    for i in range(max(args.warmup, 3)):
        tr.step(*ring[i % 4])
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # A block is EXACTLY K steps between barrier + synchronize; with a ~3 ms step a small K is a very short region, so
    # the block is repeated until >= 100 steps were timed and the median block is reported (every block max over ranks).
    blocks = max(1, -(-100 // args.steps))
    block_ms = []
    for _ in range(blocks):
        barrier()
        e0.record()
        for i in range(args.steps):
            p = tr.step(*ring[i % 4])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        block_ms.append(t.item())
    clk = clocks.stop() if clocks else None
    ms = sorted(block_ms)[len(block_ms) // 2]
    value = BATCH * world * args.steps / (ms * 1e-3)
    scal = p.scalars_dict()

    # ---- e2e: the user-facing call (CounterGanTrainer.step_auto, what train_countergan runs per batch) with pinned HOST
    # batches: every step copies x, y host->device, replays ONE graph (target / mask draw + the whole iteration) and
    # reads the step's loss scalars device->host.  Two flavours: `e2e` reads asynchronously into pinned memory and
    # consumes the values one step later (the GPU never waits for the host); `e2e_blocking` synchronises on every
    # step's read like the reference's .item() calls (trainer.py:127-129).
    hx = [r[0].cpu().pin_memory() for r in ring]
    hy = [r[1].cpu().pin_memory() for r in ring]
    e2e_steps = max(args.steps, 100)
    host_scal = [torch.empty(p.scalars.numel(), dtype=torch.float32).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    seen = []

    def e2e_iter(i, blocking):
        p_ = tr.step_auto(hx[i % 4], hy[i % 4])
        slot = i & 1
        host_scal[slot].copy_(p_.scalars, non_blocking=True)
        evs[slot].record()
        if blocking:
            evs[slot].synchronize()
            seen.append(float(host_scal[slot][1]))
        elif i >= 1:
            evs[slot ^ 1].synchronize()           # the previous step's losses are on the host now
            seen.append(float(host_scal[slot ^ 1][1]))
        return p_

    def e2e_run(blocking):
        for i in range(3):                        # warm-up (graph capture of the draw + step on first use)
            e2e_iter(i, blocking)
        barrier()
        e0.record()
        for i in range(e2e_steps):
            e2e_iter(i, blocking)
        e1.record()
        barrier()
        t_ = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return BATCH * world * e2e_steps / (t_.item() * 1e-3)

    e2e_value = e2e_run(False)
    e2e_blocking = e2e_run(True)
    assert all(v == v for v in seen), "non-finite loss read back in the e2e loop"
    h2d = hx[0].numel() * 4 + hy[0].numel() * 8
    d2h = host_scal[0].numel() * 4

    # ---- per-launcher device time inside the step (eager, CUDA events on the launching stream)
    prof, roof, detail = None, None, None
    # every rank runs the same eager steps (the step contains collectives when world > 1); rank 0 reports
    import ctypes
    tr.use_graph = False
    _lib.check(L.pcg_profile_begin())
    nprof = 3
    for i in range(nprof):
        tr.step(*ring[i % 4])
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(L.pcg_profile_end(buf, ctypes.c_size_t(len(buf))))
    tr.use_graph = tr_eager_graph
    if rank == 0:
        prof = json.loads(buf.value.decode())
        tot = sum(v["ms"] for v in prof.values())
        for v in prof.values():
            v["ms_per_step"] = v["ms"] / nprof
            v["share"] = v["ms"] / tot if tot else 0.0
            v["launches_per_step"] = v["launches"] / nprof
        pk, which = peaks()
        detail = prof
        prof = {}
        for name, v in detail.items():      # "launcher:layer.op" -> aggregate per launcher
            a = prof.setdefault(name.split(":")[0], {"ms": 0.0, "launches": 0, "ms_per_step": 0.0, "share": 0.0,
                                                      "launches_per_step": 0.0})
            for kk in a:
                a[kk] += v[kk]
        if "conv_tc64_fprop" in prof:
            # the 26 fprop + dgrad launches of the 64->64 convolution (halo-tile tcgen05 kernel)
            k = prof["conv_tc64_fprop"]
            flops_per_step = 26 * CONV_FLOPS
            achieved = flops_per_step / (k["ms_per_step"] * 1e-3) / 1e12
            peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
            traffic = None
            tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
            if os.path.exists(tp):                  # dram bytes (read + write) per launch from the committed ncu capture
                traffic = json.load(open(tp)).get("traffic_bytes_per_launch")
            # the 12 data-gradient launches whose epilogue also does a BatchNorm-backward reduction move two more
            # operand tiles per sub-tile through shared memory; the other 14 (13 forward, 1 data gradient) are the plain
            # convolution
            plain = [v for n_, v in detail.items() if n_ in ("conv_tc64_fprop:g.res.fprop", "conv_tc64_fprop:g.res.dgrad")]
            plain_ms = sum(v["ms"] for v in plain)
            plain_n = sum(v["launches"] for v in plain)
            frac_plain = (plain_n * CONV_FLOPS / (plain_ms * 1e-3) / 1e12 / peak) if plain_ms > 0 else None
            roof = {"bound": "tensor",
                    "kernel": "conv_tc64s_fprop_kernel (tcgen05 halo-tile conv 64->64, row-class stacked MMAs; 13 fprop + 13 dgrad per step)",
                    "frac_plain_conv_launches": frac_plain,
                    "plain_conv_avg_launch_us": (plain_ms / plain_n * 1e3) if plain_n else None,
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                    "algorithmic_bytes_per_launch": 2 * BATCH * 784 * 64 * 2,
                    "peak_source": f"{which} bf16_tflops_sustained (kernel timed inside the step)",
                    "avg_launch_us": k["ms"] / k["launches"] * 1e3}

    if rank == 0:
        cpu = None
        if world == 1 and not args.skip_cpu:
            ref = reference_step_rate(5, 1, budget_s=30.0)
            if ref is not None:
                rate, _, threads, done = ref
                cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "reference",
                       "sample": f"unmodified reference train_countergan (baseline/_ref), {done} iterations of B={BATCH} "
                                 "after 1 warm-up"}
            else:
                rate, _, threads = cpu_step_rate(3, 1)
                cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                       "sample": f"oracle port of trainer.py:89-132, 3 steps of B={CPU_SAMPLE_BATCH} after 1 warm-up"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": f"dp{world}",
                       "cuda_graph": tr.use_graph, "l2": "working set (~3 GB of saved activations per step) >> 126 MB L2; "
                       "4 rotating input batches"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "read": "losses copied to pinned host memory every step, consumed one step later",
                    "blocking_read_value": e2e_blocking},
            "timed_blocks": blocks, "block_ms": [round(b, 3) for b in block_ms],
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
            "kernel_breakdown_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in (prof or {}).items()},
            "kernel_detail_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(
                (detail or {}).items(), key=lambda kv: -kv[1]["ms_per_step"])[:40]} if prof else {},
            "losses": {k: round(v, 5) for k, v in scal.items()},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        tr.close()                                # captured graphs hold NCCL work: release them first
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- other BASELINE configs
# (SURVEY.md §8a a8-a16).  Same JSON line; the default workload (and the one the driver benches) stays the MNIST step.
OTHER = {
    "dcgan": dict(batch=256, name="dconv_gan/mnist DCGAN, synthetic 1x64x64 (28x28 source resized, mnist_dcgan.py:43), batch 256",
                  flops=2.243e9, bytes=None),
    # conditional WGAN-GP: one "step" = one batch of the loop (critic update with gradient penalty; every n_critic = 5th batch
    # also the generator update).  Model FLOPs per sample and batch, dense (the dilation zeros of the transposed /
    # data-gradient convolutions not counted): critic pass 70.2 MMAC; critic update = 3 forward + 3 backward (x2) + gradient
    # chain + its reverse (x2) + generator forward 196 MMAC = 1036 MMAC; generator update 728 MMAC / 5 -> 1.18 GMAC.
    "cwgan": dict(batch=128, name="conditional_gan/mnist conditional WGAN-GP (mnist_wgan_conditional.py), widths 1024/1024/1024, "
                  "synthetic 1x28x28, batch 128, n_critic 5 (4 of 5 steps critic-only)", flops=2.36e9, bytes=None),
    "kc": dict(batch=4096, name="conditional_counteRGAN/house_sales_kc_usa tabular CounteRGAN, synthetic KC-shaped features, batch 4096",
               flops=0.79e6, bytes=424.0),
    "moons_cf": dict(batch=64, name="conditional_counteRGAN/moons tabular CounteRGAN, batch 64", flops=None, bytes=None),
    "cgan_moons": dict(batch=1024, name="conditional_gan/moons class-conditional MLP GAN, batch 1024", flops=47e3, bytes=280.0),
    "simple_moons": dict(batch=128, name="simple_gan/moons MLP GAN on make_moons 2-D points, batch 128", flops=47e3, bytes=280.0),
}


def _other_setup(kind, B, dev):
    """Returns (native_step(i), cpu_step(i)) closures on identical synthetic state; the CPU side is the oracle port."""
    import torch
    from collections import OrderedDict
    import pcg_b200  # noqa: F401
    cu = lambda t: None if t is None else t.to(dev).contiguous()  # noqa: E731
    if kind == "dcgan":
        from oracle import dcgan as O
        from pcg_b200.dcgan import DcganPlan
        PG, PD = O.synth_params(O.g_shapes(), 5), O.synth_params(O.d_shapes(), 6)
        S = O.make_state(PG, O.buffers(O.g_shapes()), PD, O.buffers(O.d_shapes()))
        plan = DcganPlan(B, dev) if dev is not None else None
        if plan is not None:
            plan.G.load(PG); plan.D.load(PD); plan.refresh()
        ring = [O.synth_batch(B, 70 + i) for i in range(4)]
        dring = [tuple(cu(t) for t in b) for b in ring] if plan is not None else None
        cpu_ring = [O.synth_batch(16, 90 + i) for i in range(2)]
        return (lambda i: plan.step(*dring[i % 4])), (lambda i: O.dcgan_step(S, *cpu_ring[i % 2])), 16
    if kind == "cwgan":
        from oracle import wgan_gp as O
        from pcg_b200.wgan import Hyperparameter, WganGpPlan
        ohp = O.Hyper()
        PG, PC = O.synth_params(O.g_shapes(ohp), 7), O.synth_params(O.c_shapes(ohp), 8)
        plan = WganGpPlan(Hyperparameter(), B, dev) if dev is not None else None
        if plan is not None:
            plan.G.load(PG); plan.C.load(PC); plan.refresh()
        keys = ("real", "labels", "noise", "alpha", "labels_g", "noise_g")
        dring = [[cu(O.synth_batch(ohp, B, 70 + i)[k]) for k in keys] for i in range(4)] if plan is not None else None
        S = O.make_state(PG, O.g_buffers(ohp), PC)
        cpu_ring = [O.synth_batch(ohp, 16, 90 + i) for i in range(2)]
        eye = torch.eye(ohp.num_classes)

        def native_step(i):
            d = dring[i % 4]
            return plan.step(*d) if i % ohp.n_critic == 0 else plan.step(*d[:4])

        def cpu_step(i):
            c = cpu_ring[i % 2]
            O.critic_step(S, ohp, c["real"], eye[c["labels"]], c["noise"], c["alpha"])
            if i % ohp.n_critic == 0:
                O.generator_step(S, ohp, c["noise_g"], eye[c["labels_g"]])
        return native_step, cpu_step, 16
    if kind in ("kc", "moons_cf"):
        from oracle import tabular_countergan as T
        if kind == "kc":
            from pcg_b200.tabular.kc import KcPlan
            gs, ds, cs = T.kc_shapes()
            PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
            BD, BC = T.sn_buffers(T.kc_d_dims(), 4), T.bn_buffers(cs, 5, randomize=True)
            S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC, BC)
            nv = T.kc_norm_vals()
            plan = None
            if dev is not None:
                cat = OrderedDict((f, {"n": n, "raw_values": T.KC_RAW[f]}) for f, n in T.KC_CAT.items())
                plan = KcPlan(B, dev, cat, T.KC_CONT)
                plan.G.load(PG); plan.C.load(PC)
                for j, nm in enumerate(plan.c_bn_names):
                    plan.c_rm[j].copy_(BC[nm + ".running_mean"]); plan.c_rv[j].copy_(BC[nm + ".running_var"])
                plan.D.flat.load(PD)
                for i, L in enumerate(plan.D.layers):
                    L.u.copy_(BD[f"net.{2 * i}.weight_u"]); L.v.copy_(BD[f"net.{2 * i}.weight_v"])
                plan.refresh()
            ring = [T.kc_batch(B, 80 + i) for i in range(4)]
            dring = [(cu(x), cu(y), cu(t), cu(m), [cu(e) for e in noise]) for x, y, t, m, noise in ring] if plan else None
            return (lambda i: plan.step(*dring[i % 4])), (lambda i: T.kc_step(S, *ring[i % 4], nv)), B
        from pcg_b200.tabular.moons import MoonsPlan
        gs, ds, cs = T.moons_shapes()
        PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
        BD = T.sn_buffers(T.moons_d_dims(), 4)
        S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC)
        plan = None
        if dev is not None:
            plan = MoonsPlan(B, dev)
            plan.G.load(PG); plan.C.load(PC); plan.D.flat.load(PD)
            for i, L in enumerate(plan.D.layers):
                L.u.copy_(BD[f"net.{2 * i}.weight_u"]); L.v.copy_(BD[f"net.{2 * i}.weight_v"])
            plan.refresh()
        ring = [T.moons_batch(B, 50 + i) for i in range(4)]
        dring = [tuple(cu(t) for t in b) for b in ring] if plan else None
        return (lambda i: plan.step(*dring[i % 4])), (lambda i: T.moons_step(S, *ring[i % 4])), B
    from oracle import moons_gan as M
    from pcg_b200.moons.gan import MlpGanPlan
    label_dim = 2 if kind == "cgan_moons" else 0
    gs, ds = M.shapes(label_dim=label_dim)
    PG, PD = M.synth_params(gs, 1), M.synth_params(ds, 2)
    S = M.make_state(PG, PD)
    plan = None
    if dev is not None:
        plan = MlpGanPlan(B, 32, label_dim, 128, dev)
        plan.G.load({"net." + k: v for k, v in PG.items()}); plan.D.load({"net." + k: v for k, v in PD.items()})
        plan.refresh()
    ring = [M.synth_batch(B, 300 + i, label_dim=label_dim) for i in range(4)]
    dring = [tuple(cu(t) for t in b) for b in ring] if plan else None
    return (lambda i: plan.step(*dring[i % 4])), (lambda i: M.gan_step(S, *ring[i % 4])), B


def run_other(args):
    import torch
    spec = OTHER[args.workload]
    B = spec["batch"]
    metric = "GAN train samples/sec (G+D step) on " + args.workload
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        _, cpu_step, cb = _other_setup(args.workload, B, None)
        steps = min(max(args.steps, 1), 40)
        cpu_step(0)
        t0 = time.perf_counter()
        for i in range(steps):
            cpu_step(i)
        dt = time.perf_counter() - t0
        rate = cb * steps / dt
        print(json.dumps({"impl": "reference", "metric": metric, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": 1, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": spec["name"], "cpu_batch_per_step": cb},
                          "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                           "sample": f"oracle port, {steps} steps of B={cb}"},
                          "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: libpcg has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from pcg_b200 import _lib
    dist = None
    if world > 1 and args.workload == "dcgan":       # data parallel (the plan picks the process group up); the MLP
        import torch.distributed as dist             # configs are replicas only
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    native_step, cpu_step, cb = _other_setup(args.workload, B, dev)
    # kernels per replayed step = the launches the plan's capture(s) recorded (the eager dry run before a capture is not
    # part of a step); cwgan has two graphs, [critic + generator] and [critic only], replayed 1 : 4
    from pcg_b200 import graphs as _graphs
    _graphs.capture_log.clear()
    native_step(0)
    n1 = _lib.launch_count()
    native_step(1)
    torch.cuda.synchronize()
    log = list(_graphs.capture_log)
    if not log:                                      # a plan that launches its (single) kernel directly every step
        launches = _lib.launch_count() - n1
    else:
        launches = int(round((log[0] + 4 * log[1]) / 5)) if args.workload == "cwgan" else int(sum(log))
    for i in range(max(args.warmup, 3)):
        native_step(i)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        sc = native_step(i)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:                             # max over ranks
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    clk = clocks.stop() if clocks else None
    value = B * world * args.steps / (ms * 1e-3)      # replicas only: every rank runs the same step
    # e2e: the same step with the scalars read back every step (inputs of these plans are copied in by step())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        host = native_step(i).cpu()
    e2e = B * world * args.steps / (time.perf_counter() - t0)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if not args.skip_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_step(0)
        n = 3 if args.workload == "dcgan" else (5 if args.workload == "cwgan" else 10)
        t0 = time.perf_counter()
        for i in range(n):
            cpu_step(i)
        cpu = {"value": cb * n / (time.perf_counter() - t0), "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"oracle port, {n} steps of B={cb} after 1 warm-up"}
    pk, which = peaks()
    step_s = ms * 1e-3 / args.steps
    floor = None
    if args.workload not in ("dcgan", "cwgan") and launches > 4:
        # what bounds an operator-composed plan: the kernel -> kernel dependency latency of a CUDA graph.  Measured live:
        # the same number of (empty) launches as ONE dependency chain, and the plan's own critical path x that latency.
        from pcg_b200 import graphs, ops as K
        buf = torch.zeros(256, device=dev)
        body = lambda: [K.unary(buf, K.SCALE, buf, 1.0) for _ in range(int(launches))]    # noqa: E731
        body()
        g = graphs.capture(body)
        g.replay()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(10):
            g.replay()
        f1.record()
        torch.cuda.synchronize()
        chain_ms = f0.elapsed_time(f1) / 10
        floor = {"launches": int(launches), "one_chain_of_empty_nodes_ms": round(chain_ms, 4),
                 "us_per_dependent_node": round(chain_ms * 1e3 / launches, 3)}
        plan = next((c.cell_contents for c in (native_step.__closure__ or ()) if hasattr(c.cell_contents, "run")), None)
        prog = getattr(getattr(plan, "run", None), "program", None)
        if prog is not None:
            floor.update({"dataflow_operators": len(prog.ops), "critical_path_operators": prog.critical_path(),
                          "streams": prog.n_streams})
    if spec["flops"] and args.workload in ("dcgan", "cwgan"):
        ach = spec["flops"] * B / step_s / 1e12
        roof = {"bound": "tensor", "kernel": "whole step (64..512-channel convolutions on tcgen05, plain bf16 operands by default - PCG_TC_TERMS=3 selects bf16x3; fp32 storage; one-channel layers on CUDA cores)" if args.workload == "dcgan" else
                "whole step, model FLOPs (256..1024-channel convolutions and linear layers on the tcgen05 forward / weight-gradient kernels with bf16x3 operands = 3x the tensor work, strided data gradients as dilated forward convolutions = 4x; PCG_TC_TERMS=1 selects plain bf16)",
                "achieved": ach, "peak": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), "unit": "TFLOP/s",
                "frac": ach / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), "traffic": None}
    else:
        by = (spec["bytes"] or 0.0) * B
        ach = by / step_s / 1e9
        roof = {"bound": "hbm", "kernel": "whole step (launch/latency bound: %d kernels in one CUDA graph)" % launches,
                "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None}
    print(json.dumps({"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": ("bf16" if os.environ.get("PCG_TC_TERMS", "1") == "1" else "bf16x3 (fp32-equivalent)") if args.workload == "dcgan" else (("bf16" if os.environ.get("PCG_TC_TERMS", "3") == "1" else "bf16x3 (fp32-equivalent)") if args.workload == "cwgan" else "f32"),
                      "data": "synthetic",
                      "config": {"workload": spec["name"], "global_batch": B * world,
                                 "parallelism": ("dp%d" % world) if args.workload == "dcgan" else "replicas",
                                 "cuda_graph": True, "l2": "working set fits L2 for the MLP configs (state < 1 MB); "
                                 "4 rotating input batches"},
                      "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": int(host.numel() * 4)},
                      "gpu_launches": int(launches * args.steps), "launches_per_step": int(launches),
                      "roofline": roof, "graph_launch_floor": floor, "cpu_baseline": cpu, "clocks": clk}), flush=True)


# --------------------------------------------------------------------------------------------- widened rows
def run_widened2(args):
    """SURVEY 8f rows 2 and 3, same JSON line:
    --workload mnist_clf_train   classifier pre-training iteration (mnist/trainer.py:8-39), batch 128
    --workload kc_clf_train      KC classifier pre-training iteration (house_sales_kc_usa/trainer.py:85-96), batch 128
    --workload mnist_eval        evaluate_generator_per_target sweep (eval_utils.py:78-110): 10 targets x batch 512"""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import mnist_countergan as O
    w = args.workload
    cpu, line = None, None
    if w == "mnist_clf_train":
        from oracle import mnist_classifier as OC
        B, metric = 128, "MNIST classifier pre-training samples/sec"
        PC = O.synth_params(O.c_param_shapes(), 3, "C")
        batches = [O.synth_batch(B, 60 + i)[:2] for i in range(4)]

        def cpu_run(n):
            S = OC.make_state(PC)
            m2, m1 = OC.synth_masks(B, 1)
            OC.train_step(S, *batches[0], m2, m1)
            t0 = time.perf_counter()
            for i in range(n):
                OC.train_step(S, *batches[i % 4], m2, m1)
            return B * n / (time.perf_counter() - t0)

        def native():
            from pcg_b200.mnist.classifier_trainer import ClassifierPlan
            plan = ClassifierPlan(B, "cuda")
            plan.C.load(PC)
            plan.refresh()
            db = [(x.cuda(), y.cuda()) for x, y in batches]
            return lambda i: plan.step(*db[i % 4])
    elif w == "kc_clf_train":
        from oracle import kc_classifier as OK
        B, metric = 128, "KC classifier pre-training samples/sec"
        g = torch.Generator().manual_seed(0)
        PC = {k: (torch.randn(*s, generator=g) * 0.1 if len(s) == 2 else torch.ones(*s) if ".weight" in k else torch.zeros(*s))
              for k, s in OK.shapes().items()}
        xs = [(torch.rand(B, 17, generator=g), torch.randint(0, 4, (B,), generator=g)) for _ in range(4)]
        cw = torch.ones(4)

        def cpu_run(n):
            S = OK.make_state(PC)
            masks = OK.synth_masks(B, 1)
            OK.train_step(S, *xs[0], masks, cw)
            t0 = time.perf_counter()
            for i in range(n):
                OK.train_step(S, *xs[i % 4], masks, cw)
            return B * n / (time.perf_counter() - t0)

        def native():
            from pcg_b200.tabular.kc_classifier import KcClassifierPlan
            plan = KcClassifierPlan(B, "cuda", class_weights=cw)
            plan.C.load(PC)
            plan.refresh()
            db = [(x.cuda(), y.cuda()) for x, y in xs]
            return lambda i: plan.step(*db[i % 4])
    else:
        from oracle import mnist_eval as OE
        B, metric = 512, "MNIST per-target counterfactual evaluation samples/sec (10 targets per sample)"
        x, y, _, _ = O.synth_batch(B, 5)

        def cpu_run(n):
            S = {"G": O.synth_params(O.g_param_shapes(), 1, "G"), "GB": O.g_buffers(), "C": O.synth_params(O.c_param_shapes(), 3, "C")}
            xs_, ys_ = x[:64], y[:64]
            t0 = time.perf_counter()
            for _ in range(max(n // 10, 1)):
                OE.per_target(S, [(xs_, ys_)])
            return 64 * max(n // 10, 1) / (time.perf_counter() - t0)

        def native():
            from pcg_b200.mnist import eval_utils as EV
            from pcg_b200.mnist.models.classifier import CNNClassifier
            from pcg_b200.mnist.models.generator import ResidualGenerator
            torch.manual_seed(0)
            G, C = ResidualGenerator().cuda().eval(), CNNClassifier().cuda().eval()
            xd, yd = x.cuda(), y.cuda()

            def one(i):
                for t in range(10):
                    EV.counterfactual_metrics(G, C, xd, yd, torch.full_like(yd, t))
            return one
    unit = "samples/s"
    if not args.skip_cpu or args.impl == "reference":
        rate = cpu_run(10 if w != "mnist_eval" else 10)
        cpu = {"value": rate, "unit": unit, "cores": os.cpu_count(), "kind": "port", "sample": "oracle port, 10 iterations"}
    if args.impl == "reference":
        line = {"impl": "reference", "metric": metric, "value": cpu["value"], "unit": unit, "cpu_baseline": cpu}
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (native arm) needs a CUDA device: libpcg has no CPU fallback")
        import pcg_b200  # noqa: F401
        from pcg_b200 import _lib
        step = native()
        for i in range(max(args.warmup, 3)):
            step(i)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        line = {"metric": metric, "value": B / (ms * 1e-3), "unit": unit, "ms_per_step": ms, "n_gpus": 1, "steps": args.steps,
                "dtype": "bf16" if w == "mnist_eval" else "f32", "data": "synthetic", "higher_is_better": True,
                "config": {"workload": w, "batch": B}, "gpu_launches": int(_lib.launch_count() - n0), "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def run_widened(args):
    """--workload mnist_infer: eval-mode generator forward (BatchNorm folded), batch 512.
    --workload mnist_loader: on-device input pipeline, one shuffled epoch of 54,000 uint8 images, batch 512."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pk, which = peaks()
    if args.impl == "reference" or not torch.cuda.is_available():
        if args.impl != "reference":
            raise SystemExit("bench.py (native arm) needs a CUDA device: libpcg has no CPU fallback")
    cpu = None
    torch.set_num_threads(os.cpu_count() or 1)
    if args.workload == "mnist_infer":
        from oracle import mnist_countergan as O
        B = 512
        metric, unit = "MNIST CounteRGAN generator inference samples/sec (eval forward)", "samples/s"
        if not args.skip_cpu or args.impl == "reference":
            PG, BG = O.synth_params(O.g_param_shapes(), 1, "G"), O.g_buffers()
            xb = O.synth_batch(128, 5)
            with torch.no_grad():
                O.g_forward(PG, BG, xb[0], xb[2], xb[3], training=False)
                t0 = time.perf_counter()
                for _ in range(3):
                    O.g_forward(PG, BG, xb[0], xb[2], xb[3], training=False)
            cpu = {"value": 3 * 128 / (time.perf_counter() - t0), "unit": unit, "cores": os.cpu_count(), "kind": "port",
                   "sample": "oracle eval forward, 3 batches of 128"}
        if args.impl == "reference":
            line = {"impl": "reference", "metric": metric, "value": cpu["value"], "unit": unit, "cpu_baseline": cpu}
        else:
            import pcg_b200  # noqa: F401
            from pcg_b200.mnist.models.generator import ResidualGenerator
            torch.manual_seed(0)
            G = ResidualGenerator().cuda().eval()
            x, y, t, m = (v.cuda().contiguous() for v in O.synth_batch(B, 5))
            with torch.no_grad():
                for _ in range(max(args.warmup, 3)):
                    G(x, t, m)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    G(x, t, m)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            ach = 2 * 377.524224e6 * B / (ms * 1e-3) / 1e12
            peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
            line = {"metric": metric, "value": B / (ms * 1e-3), "unit": unit, "ms_per_step": ms, "dtype": "bf16",
                    "config": {"workload": "conditional_counteRGAN/mnist ResidualGenerator.eval() forward, batch 512"},
                    "roofline": {"bound": "tensor", "kernel": "whole forward (15 convolutions, BatchNorm folded)",
                                 "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None},
                    "cpu_baseline": cpu}
    else:
        metric, unit = "MNIST input pipeline samples/sec (shuffled epoch, batch 512)", "samples/s"
        N, B = 54000, 512
        g = torch.Generator().manual_seed(0)
        u8 = torch.randint(0, 256, (N, 28, 28), generator=g, dtype=torch.uint8)
        yv = torch.randint(0, 10, (N,), generator=g)
        if not args.skip_cpu or args.impl == "reference":
            try:
                from torchvision import transforms
                from PIL import Image
                tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,), (0.5,))])

                class DS(torch.utils.data.Dataset):
                    def __len__(self):
                        return 8192

                    def __getitem__(self, i):
                        return tf(Image.fromarray(u8[i].numpy(), mode="L")), int(yv[i])
                dl = torch.utils.data.DataLoader(DS(), batch_size=128, shuffle=True, num_workers=4)
                for _ in dl:
                    break
                t0, n = time.perf_counter(), 0
                for xb, _ in dl:
                    n += xb.shape[0]
                cpu = {"value": n / (time.perf_counter() - t0), "unit": unit, "cores": 4, "kind": "reference",
                       "sample": "data_utils.py:9-12,26 as written: torchvision transforms on PIL images, "
                                 "DataLoader(batch 128, 4 workers), 8192 images"}
            except Exception as e:      # torchvision / PIL missing
                cpu = {"value": None, "unit": unit, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
        if args.impl == "reference":
            line = {"impl": "reference", "metric": metric, "value": cpu["value"], "unit": unit, "cpu_baseline": cpu}
        else:
            import pcg_b200  # noqa: F401
            from pcg_b200.mnist import data_utils as DU
            loader = DU.DeviceLoader(u8.cuda(), yv.cuda(), None, B, shuffle=True)
            for _ in loader:
                pass
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in loader:
                pass
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            ach = N * 3936 / (ms * 1e-3) / 1e9
            line = {"metric": metric, "value": N / (ms * 1e-3), "unit": unit, "ms_per_step": ms / len(loader), "dtype": "u8->f32",
                    "config": {"workload": "54,000 resident uint8 28x28 images, shuffled epoch, batch 512"},
                    "roofline": {"bound": "hbm", "kernel": "u8_batch_kernel (launch-bound: 1 launch + index slice per batch)",
                                 "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                                 "traffic": None},
                    "cpu_baseline": cpu}
    line.update({"n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
                 "scaling": "weak", "vs_baseline": None, "data": "synthetic"})
    print(json.dumps(line), flush=True)


def flops_tc_fprop_per_step():
    """Algorithmic FLOPs executed by conv_tc_fprop launches in one step at B=512 (bf16 plan):
    G: 13 fprop + 13 dgrad of the 64->64 conv; D (3 tensor-core layers): fwd on 2B + fwd on B; C: conv.4 + fc.1."""
    g = 26 * CONV_FLOPS
    d_layers = (49 * 128 * 576, 16 * 256 * 1152, 4 * 256 * 2304)
    d = sum(2.0 * m for m in d_layers) * (3 * BATCH)
    c = 2.0 * BATCH * (49 * 128 * 576 + 6272 * 256)
    return g + d + c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PCG_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--workload", default="mnist", choices=["mnist", "mnist_infer", "mnist_loader", "mnist_clf_train", "kc_clf_train", "mnist_eval"] + sorted(OTHER))
    args = ap.parse_args()
    if args.workload in ("mnist_clf_train", "kc_clf_train", "mnist_eval"):
        run_widened2(args)
        return
    if args.workload in ("mnist_infer", "mnist_loader"):
        run_widened(args)
    elif args.workload != "mnist":
        run_other(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
