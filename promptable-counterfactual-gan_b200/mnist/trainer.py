"""Mirror of ``conditional_counteRGAN/mnist/trainer.py`` for the GAN hot path.

``train_countergan(generator, discriminator, classifier, train_loader, cfg, device)`` has the
reference's signature (trainer.py:76) and side effects (prints, ``gan_losses.png`` when matplotlib is
present, ``torch.save(generator.state_dict(), cfg.generator_path)``), but every iteration is ONE native
step (libpcg) instead of ~1.4 k torch ops: generator forward, D step, G step and both Adam updates.

The modules may be the mirror classes of ``pcg_b200.mnist.models`` or the reference's own classes: the
trainer adopts their parameters into flat arenas (``binding.bind``), so ``state_dict()`` stays valid.

Data parallel: if ``torch.distributed`` is initialised with world_size > 1 the two gradient arenas are
all-reduced (NCCL) at the two points the algorithm requires — D's before ``opt_d.step()`` because the
G step uses the updated D (trainer.py:112 -> :116), G's before ``opt_g.step()``.
"""
import os

import torch

from .. import graphs
from .. import ops as K
from . import binding
from . import plan as P


def grad_norm(parameters):
    """trainer.py:41-42."""
    return torch.sqrt(sum((p.grad.data.norm() ** 2) for p in parameters if p.grad is not None)).item()


class _Rng:
    """Per-device Philox state of the native mask / target kernel: three device uint64 {stream offset, ticket, key}.
    The key follows torch's CUDA generator seed and ``torch.manual_seed`` restarts the stream (as it would restart the
    reference's draws): a re-seed is recognised by the generator's seed changing or its offset falling back below the
    mark this class left on it.  The state lives in device memory and is reset IN PLACE, so a launch captured in a CUDA
    graph draws fresh numbers at every replay and follows a re-seed without being re-captured."""
    _states = {}

    @classmethod
    def get(cls, device):
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        seed, gen, off = torch.initial_seed(), None, 0
        try:
            gen = torch.cuda.default_generators[device.index]
            seed, off = gen.initial_seed(), gen.get_offset()
        except Exception:                                   # generator offsets unavailable: seed changes only
            gen = None
        st = cls._states.get(device)
        if st is None or st["seed"] != seed or off < st["mark"]:
            rank = 0
            dist, _ = _world()
            if dist is not None:
                rank = dist.get_rank()                     # per-rank sub-streams
            key = (seed + 0x9E3779B97F4A7C15 * rank) & (2 ** 63 - 1)
            if st is None:
                st = cls._states[device] = {"state": torch.zeros(3, dtype=torch.int64, device=device)}
            st["state"].copy_(torch.tensor([0, 0, key], dtype=torch.int64))
            st["seed"] = seed
            if gen is not None:
                off = (off + 4) // 4 * 4
                gen.set_offset(off)                         # leave a mark: manual_seed() puts the offset back to 0
            st["mark"] = off
        return 0, st["state"]


def build_mask(x, patch_size, device, num_modifiable_patches=None):
    """Random binary patch mask with the distribution of trainer.py:45-72 - per sample a uniformly random subset of
    ``num_modifiable_patches`` patches (fair coins per patch when that is None or >= the patch count), nearest-upsampled
    and repeated over the channels - in ONE native launch (``pcg_build_mask``, Philox) instead of the reference's Python
    loop of ``bs`` randperm calls."""
    bs, c, h, w = x.shape
    if torch.device(device).type != "cuda":
        raise RuntimeError("pcg_b200.mnist.trainer.build_mask needs a CUDA device (there is no CPU fallback)")
    mask = torch.empty(bs, c, h, w, dtype=torch.float32, device=device)
    seed, state = _Rng.get(device)
    with torch.cuda.device(mask.device):
        K.build_mask(bs, c, h, w, patch_size, num_modifiable_patches, mask, None, seed=seed, rng_state=state)
    return mask


def draw_target_and_mask(x, cfg, device):
    """trainer.py:94-95 (``target_y = torch.randint(0, num_classes, (bs,))`` and ``build_mask``) as one launch."""
    bs, c, h, w = x.shape
    mask = torch.empty(bs, c, h, w, dtype=torch.float32, device=device)
    target = torch.empty(bs, dtype=torch.int64, device=device)
    seed, state = _Rng.get(device)
    with torch.cuda.device(mask.device):
        K.build_mask(bs, c, h, w, cfg.patch_size, cfg.num_modifiable_patches, mask, target, cfg.num_classes, seed=seed,
                     rng_state=state)
    return target, mask


_NATIVE_DRAW = draw_target_and_mask


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_world_size()
    return None, 1


class CounterGanTrainer:
    """Owns the arenas, Adam state and one native plan per batch size; ``step`` runs one iteration."""

    def __init__(self, generator, discriminator, classifier, cfg, device, precision=None, use_graph=True):
        self.G, self.D, self.C, self.cfg = generator, discriminator, classifier, cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pcg_b200.mnist.trainer needs a CUDA device (there is no CPU fallback)")
        self.ga = binding.bind(generator, 0, self.device)
        self.da = binding.bind(discriminator, 1, self.device)
        self.ca = binding.bind(classifier, 2, self.device)
        self.base_ch, self.n_res = generator._pcg_dims
        self.running, self.nbt = generator._pcg_bn
        self.dist, self.world = _world()
        self.precision = precision or os.environ.get("PCG_PRECISION", "bf16")
        self.use_graph = use_graph and os.environ.get("PCG_NO_GRAPH", "0") != "1"
        self.adam = None
        self.plans = {}
        self.graphs = {}
        self.static = {}
        self._last_bs = None
        self._dp_graphs = []
        self._side = torch.cuda.Stream(device=self.device) if self.dist is not None else None

    def close(self):
        """Drops the captured graphs (they hold NCCL work captured on the process group's stream: release them before
        ``destroy_process_group``) and the native plans."""
        self.graphs.clear()
        self._dp_graphs.clear()
        for p in self.plans.values():
            p.close()
        self.plans.clear()
        import gc
        gc.collect()

    def _step_cfg_grad_scale(self):
        """Gradients are summed by the all-reduce; the 1/world_size average is folded into Adam."""
        return 1.0 / self.world

    def _step_cfg(self):
        c = self.cfg
        return P.StepConfig(g_lr=c.g_lr, d_lr=c.d_lr, lambda_adv=c.lambda_adv, lambda_cls=c.lambda_cls,
                            lambda_reg=c.lambda_reg, lambda_mask=c.lambda_mask, grad_scale=self._step_cfg_grad_scale(),
                            precision=self.precision)

    def plan(self, bs):
        p = self.plans.get(bs)
        if p is None:
            p = P.MnistStepPlan(bs, self.ga, self.da, self.ca, self.running, self.nbt, self._step_cfg(),
                                self.base_ch, self.n_res, adam_state=self.adam)
            self.adam = p.adam          # every plan shares one Adam state
            self.plans[bs] = p
        return p

    def _allreduce(self, t):
        if self.dist is not None:
            self.dist.all_reduce(t)

    # ---- data parallel: the two gradient reductions the algorithm requires.  Three overlap schemes are
    # implemented and were measured at 2 x B200 (profiles/exp_dp_overlap_r2.md); none moved the step time (3.05-3.07 ms
    # against 2.97 ms on one GPU): the messages (3.9 MB and 2.0 MB) are latency-bound, so what is exposed is two NCCL
    # latencies whatever is scheduled beside them.  They stay as switches, default off:
    #   PCG_DP_DEFER_CBWD=1  D's reduction runs beside the frozen classifier's input-gradient chain (own stream);
    #   PCG_DP_SPLIT=1       G's reduction in two buckets: [resblocks.k .. conv_out], the tail of the flat arena that the
    #                        backward pass finishes first, is reduced while blocks k-1 .. 0 are still differentiated;
    #   PCG_DP_ONE_GRAPH=1   the whole iteration including the NCCL calls captured as ONE graph (close() before
    #                        destroy_process_group).
    def _dp_split(self):
        """(residual block k, arena offset of its first parameter), or None when the generator is too shallow."""
        if self.n_res < 2 or os.environ.get("PCG_DP_SPLIT", "0") != "1":
            return None
        k = self.n_res // 2
        return k, self.ga.slots[3 + 8 * k][0]          # parameters(): embed, conv_in w/b, then 8 tensors per block

    def _dp_step(self, p, seg):
        """seg: dict of callables 'd_grads', 'c_bwd', 'mid1' (D update + G backward part 1), 'mid2' (part 2), 'g_update'
        (eager phase calls or graph replays)."""
        on_gpu = getattr(self, "device", None) is not None and self.device.type == "cuda"
        seg["d_grads"]()
        if on_gpu:
            main, side = torch.cuda.current_stream(), self._side_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                seg["c_bwd"]()
        else:                                          # host-logic tests (gloo on CPU tensors): no streams
            seg["c_bwd"]()
        wd = self.dist.all_reduce(self.da.grad, async_op=True)
        wd.wait()
        if on_gpu:
            main.wait_stream(side)
        split = self._dp_split()
        if split is None:
            seg["mid1"]()
            self.dist.all_reduce(self.ga.grad)
        else:
            off = split[1]
            seg["mid1"]()
            w_tail = self.dist.all_reduce(self.ga.grad[off:], async_op=True)
            seg["mid2"]()
            w_head = self.dist.all_reduce(self.ga.grad[:off], async_op=True)
            w_tail.wait()
            w_head.wait()
        seg["g_update"]()

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _dp_segments(self, p, x, y, t, m):
        split = self._dp_split()
        defer = os.environ.get("PCG_DP_DEFER_CBWD", "0") == "1"
        p.set_defer_c_bwd(defer)
        if split is None:
            mid1 = lambda: (p.step_d_update(), p.step_g_grads(x, y, t, m))    # noqa: E731
            mid2 = None
        else:
            k = split[0]
            mid1 = lambda: (p.step_d_update(), p.step_g_grads_part(x, y, t, m, 1, k))    # noqa: E731
            mid2 = lambda: p.step_g_grads_part(x, y, t, m, 2, k)    # noqa: E731
        return {"d_grads": lambda: p.step_d_grads(x, y, t, m), "c_bwd": p.step_c_bwd if defer else (lambda: None),
                "mid1": mid1, "mid2": mid2, "g_update": p.step_g_update}

    def _run_phases(self, p, x, y, t, m):
        if self.dist is None:
            p.step(x, y, t, m)
        else:
            self._dp_step(p, self._dp_segments(p, x, y, t, m))

    def step(self, x, y, target, mask):
        """One iteration on device tensors; returns the plan whose ``scalars`` hold the losses."""
        bs = x.shape[0]
        p = self._prepare(bs)
        if not self.use_graph:
            self._run_phases(p, x, y, target, mask)
            return p
        st = self.static.get(bs)
        if st is None:
            st = tuple(torch.empty_like(v) for v in (x, y, target, mask))
            self.static[bs] = st
        for dst, src in zip(st, (x, y, target, mask)):
            dst.copy_(src, non_blocking=True)
        g = self.graphs.get(bs)
        if g is None:
            # one eager iteration would consume a real batch; capture directly after a dry launch
            # of the attribute-setting paths on a side stream (the graph body is pure kernel launches)
            g = self._capture(p, st)
            self.graphs[bs] = g
        g()
        return p

    def _prepare(self, bs):
        """The plan of this batch size, with its packed weights and the modules' forward plans made consistent."""
        p = self.plan(bs)
        if self._last_bs is not None and self._last_bs != bs:
            # every plan owns packed copies of the conv weights, refreshed by ITS OWN Adam phases only: the loader has
            # no drop_last (data_utils.py:27; 54000 % 128 = 112), so the tail-batch plan and the full-batch plan
            # alternate and each must re-pack what the other one updated before its forward runs
            p.refresh_weights()
        self._last_bs = bs
        for m in (self.G, self.D, self.C):
            # forward-only plans cached on the modules (models/_native.py) re-pack on their next call: the native Adam
            # writes below do not bump torch's parameter version counters
            m.__dict__["_pcg_seen"] = None
        return p

    def step_auto(self, x, y):
        """One iteration as the reference loop runs it: x, y may be (pinned) host or device tensors; the target classes
        and the patch mask (trainer.py:94-95) are drawn on the device by the first node of the replayed graph."""
        bs = x.shape[0]
        if not self.use_graph or draw_target_and_mask is not _NATIVE_DRAW:
            # eager mode, or a caller replaced the draw (tests inject the reference's draws): explicit inputs
            xd = x.to(self.device, non_blocking=True).float().contiguous()
            yd = y.to(self.device, non_blocking=True).long().contiguous()
            target, mask = draw_target_and_mask(xd, self.cfg, self.device)
            return self.step(xd, yd, target, mask)
        p = self._prepare(bs)
        st = self.static.get(bs)
        if st is None:
            st = (torch.empty(bs, 1, 28, 28, device=self.device), torch.empty(bs, dtype=torch.int64, device=self.device),
                  torch.empty(bs, dtype=torch.int64, device=self.device), torch.empty(bs, 1, 28, 28, device=self.device))
            self.static[bs] = st
        st[0].copy_(x.reshape(bs, 1, 28, 28), non_blocking=True)
        st[1].copy_(y, non_blocking=True)
        seed, state = _Rng.get(self.device)          # every step: a torch.manual_seed() since the last one resets the state
        g = self.graphs.get((bs, "auto"))
        if g is None:
            cfg = self.cfg

            def draw():
                K.build_mask(bs, 1, 28, 28, cfg.patch_size, cfg.num_modifiable_patches, st[3], st[2], cfg.num_classes,
                             seed=seed, rng_state=state)
            g = self._capture(p, st, draw)
            self.graphs[(bs, "auto")] = g
        g()
        return p

    def _capture(self, p, st, pre=None):
        if self.dist is not None:
            seg = self._dp_segments(p, *st)
            d_grads = seg["d_grads"]
            seg["d_grads"] = lambda: ((pre() if pre else None), d_grads())
            if os.environ.get("PCG_DP_ONE_GRAPH", "0") == "1" and self.dist.get_backend() == "nccl":
                # the whole iteration INCLUDING the NCCL all-reduces as one graph: no host round trip and no idle gap
                # at the five segment boundaries (the collectives are captured on NCCL's own stream, forked from and
                # joined to the capture stream by the process group).  close() must run before destroy_process_group.
                g = self._capture_fn(lambda: self._dp_step(p, seg))
                self._dp_graphs.append(g)
                return g.replay
            # otherwise (gloo, or PCG_DP_ONE_GRAPH=0): capture the kernel-only segments, replay them around the reductions
            graphs_ = {k: self._capture_fn(fn) for k, fn in seg.items() if fn is not None}
            replay = {k: (graphs_[k].replay if k in graphs_ else None) for k in seg}
            return lambda: self._dp_step(p, replay)
        g = self._capture_fn(lambda: ((pre() if pre else None), p.step(*st)))
        return g.replay

    @staticmethod
    def _capture_fn(fn):
        # thread_local capture: other threads (e.g. the NCCL watchdog polling events) must not invalidate it
        return graphs.capture(fn)


def _warm_plan(tr, bs):
    """Runs the attribute-setting first launches (cudaFuncSetAttribute, tensor-map entry points) on
    throw-away state so that graph capture later sees pure kernel launches and no real batch is
    consumed: parameters, Adam state and BN buffers are snapshotted and restored."""
    p = tr.plan(bs)
    snap = [t.clone() for t in (tr.ga.data, tr.da.data, tr.running, tr.nbt, p.adam["g_m"], p.adam["g_v"],
                                p.adam["d_m"], p.adam["d_v"], p.adam["g_step"], p.adam["d_step"])]
    x = torch.zeros(bs, 1, 28, 28, device=tr.device)
    y = torch.zeros(bs, dtype=torch.int64, device=tr.device)
    p.step(x, y, y, torch.ones_like(x))
    torch.cuda.synchronize()
    for dst, src in zip((tr.ga.data, tr.da.data, tr.running, tr.nbt, p.adam["g_m"], p.adam["g_v"], p.adam["d_m"],
                         p.adam["d_v"], p.adam["g_step"], p.adam["d_step"]), snap):
        dst.copy_(src)
    p.refresh_weights()
    torch.cuda.synchronize()


def train_countergan(generator, discriminator, classifier, train_loader, cfg, device):
    """Drop-in for trainer.py:76-163."""
    tr = CounterGanTrainer(generator, discriminator, classifier, cfg, device)
    g_losses, d_losses, g_cls_losses = [], [], []
    warmed = set()

    for epoch in range(cfg.num_epochs_gan):
        acc = torch.zeros(P.NSCALARS, device=tr.device)      # device-side running sums, no per-step sync
        num_batches = 0
        p = None
        for batch_idx, (x, y) in enumerate(train_loader):
            bs = x.size(0)
            num_batches += 1
            if bs not in warmed:
                _warm_plan(tr, bs)
                warmed.add(bs)
            p = tr.step_auto(x, y)            # target / mask draw (trainer.py:94-95) + the whole iteration: one replay
            acc += p.scalars
            if batch_idx % 100 == 0:
                s = p.scalars_dict()                                                        # one sync / 100 steps
                print(f"[Epoch {epoch+1}/{cfg.num_epochs_gan}] batch {batch_idx} :: "
                      f"D(real)={s['d_real_p']:.3f}, D(fake)={s['d_fake_p']:.3f}, g_adv={s['g_adv']:.4f}, "
                      f"g_cls={s['g_cls']:.4f}, reg={s['reg_l1']:.6f}, residual_mean={s['reg_l1']:.4f}")
        tot = acc.tolist()
        n = max(num_batches, 1)
        g_losses.append(tot[1] / n)
        d_losses.append(tot[0] / n)
        g_cls_losses.append(tot[3] / n)
        # the arenas hold the all-reduced SUM of the per-rank gradients; the reference logs the norm of the averaged ones
        g_grad_norm = tr.ga.grad.norm().item() / tr.world
        d_grad_norm = tr.da.grad.norm().item() / tr.world
        print(f"[GAN] Epoch {epoch+1}/{cfg.num_epochs_gan} | "
              f"G: {g_losses[-1]:.4f}, D: {d_losses[-1]:.4f}, "
              f"G_cls: {g_cls_losses[-1]:.4f}, G_grad: {g_grad_norm:.4f}, D_grad: {d_grad_norm:.4f}")

    save_path = os.path.join(cfg.save_dir, "gan_losses.png")
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:       # matplotlib is optional here (absent in the build image)
        plt = None
        print("matplotlib unavailable: skipped the loss-curve plot")
    if plt is not None:
        plt.figure(figsize=(8, 6))
        plt.plot(g_losses, label="Generator Loss")
        plt.plot(d_losses, label="Discriminator Loss")
        plt.plot(g_cls_losses, label="Classifier Loss (g_cls)", linestyle="--")
        plt.xlabel("Epoch")
        plt.ylabel("Loss")
        plt.legend()
        plt.title("CounterGAN Losses")
        plt.savefig(save_path)
        plt.close()
        print(f"Saved GAN loss curves to {save_path}")

    torch.save(generator.state_dict(), cfg.generator_path)
    print(f"Generator saved to {cfg.generator_path}")
    return {"g_losses": g_losses, "d_losses": d_losses, "g_cls_losses": g_cls_losses}
