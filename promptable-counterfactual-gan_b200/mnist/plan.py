"""Python owner of a native MNIST CounteRGAN step plan (``pcg_mnist_plan`` in include/pcg.h).

PyTorch supplies device memory, streams and (for N > 1) ``torch.distributed``; every kernel of the
step lives in libpcg.so.  Parameters of the three networks live in flat fp32 arenas whose slices are
what the ``nn.Module`` parameters point at, so ``state_dict()`` / ``load_state_dict()`` keep working
and Adam / the gradient all-reduce see one contiguous buffer per network.
"""
import ctypes
from dataclasses import dataclass

import torch

from .. import _lib
from .. import graphs

PCG_F32, PCG_BF16 = 0, 1
SCALAR_NAMES = ["d_loss", "g_loss", "g_adv", "g_cls", "reg_l1", "mask_pen", "d_real_p", "d_fake_p",
                "d_loss_real", "d_loss_fake"]
NSCALARS = 16


class _Config(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int), ("base_ch", ctypes.c_int), ("n_resblocks", ctypes.c_int),
                ("precision", ctypes.c_int), ("g_lr", ctypes.c_float), ("d_lr", ctypes.c_float),
                ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("adam_eps", ctypes.c_float),
                ("lambda_adv", ctypes.c_float), ("lambda_cls", ctypes.c_float), ("lambda_reg", ctypes.c_float),
                ("lambda_mask", ctypes.c_float), ("residual_scaling", ctypes.c_float),
                ("grad_scale", ctypes.c_float), ("use_tensor_cores", ctypes.c_int),
                ("pollute_d_grads", ctypes.c_int)]


class _Buffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("g_params", "g_grads", "g_adam_m", "g_adam_v", "g_step", "g_bn_running", "g_bn_nbt",
                 "d_params", "d_grads", "d_adam_m", "d_adam_v", "d_step", "c_params")]


class _Inputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("x", "y", "target", "mask")]


@dataclass
class StepConfig:
    """Hyper-parameters read by the reference loop (config.py:10-15, trainer.py:77-78)."""
    g_lr: float = 5e-5
    d_lr: float = 1e-5
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    lambda_adv: float = 1.0
    lambda_cls: float = 1.0
    lambda_reg: float = 2.5
    lambda_mask: float = 2.0
    residual_scaling: float = 0.1
    grad_scale: float = 1.0
    pollute_d_grads: bool = False
    use_tensor_cores: bool = True    # bf16 only; False = same storage, CUDA-core convolutions (A/B check)
    precision: str = "bf16"      # "bf16" (tcgen05 tensor cores) or "fp32" (CUDA cores, exact mode)


def layout(net, base_ch=64, n_resblocks=6):
    """[(offset, numel)] of each parameter tensor in the flat arena + arena size (floats)."""
    L = _lib.load()
    tot = ctypes.c_longlong()
    n = L.pcg_mnist_layout(net, base_ch, n_resblocks, -1, None, None, ctypes.byref(tot))
    if n < 0:
        raise _lib.PcgError(L.pcg_last_error().decode())
    out = []
    for i in range(n):
        off, num = ctypes.c_longlong(), ctypes.c_longlong()
        L.pcg_mnist_layout(net, base_ch, n_resblocks, i, ctypes.byref(off), ctypes.byref(num), None)
        out.append((off.value, num.value))
    return out, tot.value


class Arena:
    """Flat fp32 CUDA buffer holding every parameter of one network in ``parameters()`` order."""

    def __init__(self, net, base_ch, n_resblocks, device):
        self.net, self.slots = net, None
        self.slots, self.size = layout(net, base_ch, n_resblocks)
        self.data = torch.zeros(self.size, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.size, dtype=torch.float32, device=device)

    def view(self, i, shape, grad=False):
        off, n = self.slots[i]
        return (self.grad if grad else self.data)[off:off + n].view(shape)

    def adopt(self, params):
        """Copies the tensors in and re-points ``p.data`` / ``p.grad`` at the arena slices."""
        params = list(params)
        assert len(params) == len(self.slots), (len(params), len(self.slots))
        for i, p in enumerate(params):
            off, n = self.slots[i]
            assert p.numel() == n, (i, tuple(p.shape), n)
            v = self.view(i, p.shape)
            v.copy_(p.detach().to(v.device, torch.float32))
            p.data = v
            if p.requires_grad:
                p.grad = self.view(i, p.shape, grad=True)
        return self

    def load_dict(self, tensors):
        for i, t in enumerate(tensors):
            self.view(i, t.shape).copy_(t)
        return self


class MnistStepPlan:
    """One native plan: fixed batch size, borrowed arenas, owned workspaces."""

    def __init__(self, batch, g_arena, d_arena, c_arena, bn_running, bn_nbt, cfg: StepConfig = None,
                 base_ch=64, n_resblocks=6, adam_state=None):
        if not torch.cuda.is_available():
            raise _lib.PcgError("pcg_b200 needs a CUDA device: there is no CPU fallback")
        self.L = _lib.load()
        self.cfg = cfg or StepConfig()
        self.batch, self.base_ch, self.n_resblocks = batch, base_ch, n_resblocks
        dev = g_arena.data.device
        self.device = dev
        self.g, self.d, self.c = g_arena, d_arena, c_arena
        self.bn_running, self.bn_nbt = bn_running, bn_nbt
        if adam_state is None:
            adam_state = {k: torch.zeros_like(a.data) for k, a in
                          (("g_m", g_arena), ("g_v", g_arena), ("d_m", d_arena), ("d_v", d_arena))}
            adam_state["g_step"] = torch.zeros(1, dtype=torch.int32, device=dev)
            adam_state["d_step"] = torch.zeros(1, dtype=torch.int32, device=dev)
        self.adam = adam_state
        self.scalars = torch.zeros(NSCALARS, dtype=torch.float32, device=dev)
        c = self.cfg
        self._cfg = _Config(batch, base_ch, n_resblocks, PCG_BF16 if c.precision == "bf16" else PCG_F32,
                            c.g_lr, c.d_lr, c.beta1, c.beta2, c.adam_eps, c.lambda_adv, c.lambda_cls,
                            c.lambda_reg, c.lambda_mask, c.residual_scaling, c.grad_scale,
                            1 if c.use_tensor_cores else 0, 1 if c.pollute_d_grads else 0)
        P = lambda t: t.data_ptr()  # noqa: E731
        self._buf = _Buffers(P(g_arena.data), P(g_arena.grad), P(self.adam["g_m"]), P(self.adam["g_v"]),
                             P(self.adam["g_step"]), P(bn_running), P(bn_nbt), P(d_arena.data), P(d_arena.grad),
                             P(self.adam["d_m"]), P(self.adam["d_v"]), P(self.adam["d_step"]), P(c_arena.data))
        self._plan = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self.L.pcg_mnist_plan_create(ctypes.byref(self._cfg), ctypes.byref(self._buf),
                                                    ctypes.byref(self._plan)))
        self._keep = None

    def close(self):
        if getattr(self, "_plan", None) is not None and self._plan.value:
            handle, L = self._plan, self.L
            self._plan = ctypes.c_void_p()
            # cudaFree / cudaStreamDestroy are illegal while this thread captures a graph (a finaliser can run then)
            graphs.defer_destroy(lambda: L.pcg_mnist_plan_destroy(handle))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _inputs(self, x, y, target, mask):
        for t, dt in ((x, torch.float32), (y, torch.int64), (target, torch.int64), (mask, torch.float32)):
            assert t.is_cuda and t.dtype == dt and t.is_contiguous(), (t.device, t.dtype)
        assert x.numel() == self.batch * 784 and mask.numel() == self.batch * 784
        self._keep = (x, y, target, mask)
        return _Inputs(x.data_ptr(), y.data_ptr(), target.data_ptr(), mask.data_ptr())

    def refresh_weights(self):
        _lib.check(self.L.pcg_mnist_refresh_weights(self._plan, _lib.stream_ptr()))

    # ------------------------------------------------------------------ step
    def step(self, x, y, target, mask):
        inp = self._inputs(x, y, target, mask)
        _lib.check(self.L.pcg_mnist_step(self._plan, ctypes.byref(inp), _lib.ptr(self.scalars), _lib.stream_ptr()))
        return self.scalars

    def step_d_grads(self, x, y, target, mask):
        inp = self._inputs(x, y, target, mask)
        _lib.check(self.L.pcg_mnist_step_d_grads(self._plan, ctypes.byref(inp), _lib.ptr(self.scalars),
                                                 _lib.stream_ptr()))

    def step_d_update(self):
        _lib.check(self.L.pcg_mnist_step_d_update(self._plan, _lib.stream_ptr()))

    def step_g_grads(self, x, y, target, mask):
        inp = self._inputs(x, y, target, mask)
        _lib.check(self.L.pcg_mnist_step_g_grads(self._plan, ctypes.byref(inp), _lib.ptr(self.scalars),
                                                 _lib.stream_ptr()))

    # data-parallel refinements (include/pcg.h): classifier input gradient as its own phase, generator backward in two parts
    def set_defer_c_bwd(self, on):
        _lib.check(self.L.pcg_mnist_set_defer_c_bwd(self._plan, 1 if on else 0))

    def step_c_bwd(self):
        _lib.check(self.L.pcg_mnist_step_c_bwd(self._plan, _lib.stream_ptr()))

    def step_g_grads_part(self, x, y, target, mask, part, split):
        inp = self._inputs(x, y, target, mask)
        _lib.check(self.L.pcg_mnist_step_g_grads_part(self._plan, ctypes.byref(inp), _lib.ptr(self.scalars), part, split,
                                                      _lib.stream_ptr()))

    def step_g_update(self):
        _lib.check(self.L.pcg_mnist_step_g_update(self._plan, _lib.stream_ptr()))

    def scalars_dict(self):
        v = self.scalars.tolist()
        return {n: v[i] for i, n in enumerate(SCALAR_NAMES)}

    # ------------------------------------------------------------------ forwards
    def g_forward(self, x, target, mask, training):
        raw = torch.empty(self.batch, 1, 28, 28, dtype=torch.float32, device=self.device)
        masked = torch.empty_like(raw)
        _lib.check(self.L.pcg_mnist_g_forward(self._plan, _lib.ptr(x), _lib.ptr(target), _lib.ptr(mask),
                                              1 if training else 0, _lib.ptr(raw), _lib.ptr(masked),
                                              _lib.stream_ptr()))
        return raw, masked

    def d_forward(self, x, cond):
        out = torch.empty(self.batch, 1, dtype=torch.float32, device=self.device)
        _lib.check(self.L.pcg_mnist_d_forward(self._plan, _lib.ptr(x), _lib.ptr(cond), _lib.ptr(out),
                                              _lib.stream_ptr()))
        return out

    def c_forward(self, x):
        out = torch.empty(self.batch, 10, dtype=torch.float32, device=self.device)
        _lib.check(self.L.pcg_mnist_c_forward(self._plan, _lib.ptr(x), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def debug_tensor(self, name):
        """Copy of an internal NHWC tensor as fp32 (flat)."""
        p, n, dt = ctypes.c_void_p(), ctypes.c_longlong(), ctypes.c_int()
        _lib.check(self.L.pcg_mnist_debug_tensor(self._plan, name.encode(), ctypes.byref(p), ctypes.byref(n),
                                                 ctypes.byref(dt)))
        out = torch.empty(n.value, dtype=torch.bfloat16 if dt.value == PCG_BF16 else torch.float32,
                          device=self.device)
        _lib.check(self.L.pcg_memcpy_d2d(_lib.ptr(out), p, ctypes.c_size_t(n.value * out.element_size()),
                                         _lib.stream_ptr()))
        torch.cuda.synchronize()
        return out.float()
