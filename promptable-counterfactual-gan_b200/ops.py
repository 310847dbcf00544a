"""Thin ctypes wrappers of libpcg's primitive operators (include/pcg.h, "Primitive operators").

Every function enqueues one or a few kernels on the current CUDA stream and returns nothing; outputs are
caller-allocated torch tensors (fp32, contiguous).  No torch arithmetic happens here — the step plans built on
these wrappers are captured in a CUDA graph, so the Python overhead is paid once.
"""
import ctypes
import os

import torch

from . import _lib

RELU, LRELU, SIGMOID, TANH, SCALE, COPY = 1, 2, 3, 4, 5, 6
ACT_NONE, ACT_LRELU, ACT_RELU = 0, 1, 2
ADD, MUL = 0, 1
GAN_LOG, GAN_BCE, GAN_WASSERSTEIN = 0, 1, 2

_f = ctypes.c_float
_i = ctypes.c_int
_ll = ctypes.c_longlong
P = _lib.ptr


def _L():
    return _lib.load()


# ---------------------------------------------------------------- dataflow recording (pcg_b200.dataflow)
# Every operator wrapper below is declared with the names of the arguments it WRITES; while a recorder is installed the
# call is not launched but appended to it with its read set (every tensor reachable from the arguments) and write set,
# so that the step plans can be re-emitted onto several streams with exactly the dependencies the data flow requires.
_rec = None
_cache_on = False       # tensor-core operand cache (set_operand_cache): the wrappers report every operator's writes to it


def _tensors(v, out):
    if v is None:
        return out
    if isinstance(v, torch.Tensor):
        out.append(v)
    elif isinstance(v, tuple) and len(v) == 3 and isinstance(v[0], torch.Tensor) and isinstance(v[1], int):
        out.append(v)                              # (tensor, first column, end column): a column window of a 2-D tensor
    elif isinstance(v, (list, tuple)):
        for e in v:
            _tensors(e, out)
    elif isinstance(v, BNState):
        for e in vars(v).values():
            _tensors(e, out)
    return out


def _op(*writes, reads=None):
    """Decorator: ``writes`` are argument names, or callables ``bound_arguments -> iterable of tensors``; ``reads``
    (optional callable) replaces the default read set, every tensor reachable from the arguments."""
    import functools
    import inspect

    def deco(fn):
        sig = inspect.signature(fn)

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            if _rec is None and not _cache_on:
                return fn(*args, **kwargs)
            b = sig.bind(*args, **kwargs)
            b.apply_defaults()
            w = []
            for name in writes:
                _tensors(name(b.arguments) if callable(name) else b.arguments[name], w)
            if _rec is None:
                # cached bf16 conversions of the tensors this operator overwrites are stale from here on
                inval = _L().pcg_operand_cache_invalidate
                for t in w:
                    t = t[0] if isinstance(t, tuple) else t
                    inval(ctypes.c_void_p(t.data_ptr()), _ll(t.numel() * t.element_size()))
                return fn(*args, **kwargs)
            r = _tensors(reads(b.arguments) if reads is not None else list(b.arguments.values()), [])
            _rec.add(fn, args, kwargs, r, w)
            return None
        wrapper.__wrapped_op__ = fn
        return wrapper
    return deco


def _s():
    return _lib.stream_ptr()


def _chk(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.is_contiguous(), "libpcg operators need contiguous CUDA tensors"


def stat_scratch(C, device):
    L = _L()
    L.pcg_stat_scratch_floats.restype = ctypes.c_longlong
    return torch.zeros(int(L.pcg_stat_scratch_floats(C)), device=device)


def set_conv_tensor_cores(on):
    """Routes eligible conv_fprop / conv_dgrad / conv_wgrad calls to the tcgen05 kernels (bf16 operands, fp32
    accumulation); off = exact fp32 on the CUDA cores.  See include/pcg.h pcg_set_conv_tensor_cores."""
    _lib.check(_L().pcg_set_conv_tensor_cores(1 if on else 0))


def set_operand_cache(on):
    """Tensor-core operand cache (include/pcg.h): on = reuse the bf16 conversions of unchanged operands.  While it is on,
    every operator wrapper of this module invalidates the conversions of the tensors it writes; tensors written by anything
    else (torch copies of the step inputs) need ``operand_cache_clear()``.  Returns the previous setting."""
    global _cache_on
    L = _L()
    L.pcg_operand_cache_invalidate.restype = None
    L.pcg_operand_cache_clear.restype = None
    prev = bool(L.pcg_set_operand_cache(1 if on else 0))
    _cache_on = bool(on)
    return prev


def operand_cache_clear():
    L = _L()
    L.pcg_operand_cache_clear.restype = None
    L.pcg_operand_cache_clear()


def set_conv_tensor_core_terms(terms):
    """3 = bf16x3 operands (fp32-equivalent), 1 = plain bf16 operands; returns the previous setting."""
    return int(_L().pcg_set_conv_tensor_core_terms(int(terms)))


# ---------------------------------------------------------------- convolution / linear
@_op("wf", "wd")
def pack_weights(w, k, wf=None, wd=None, perm_hw=0):
    """torch OIHW (or [out,in] with k=1) -> wf [Cout][k*k][Cin], wd [Cin][k*k][Cout]."""
    Cout, Cin = w.shape[0], w.shape[1]
    _chk(w, wf, wd)
    _lib.check(_L().pcg_pack_conv_weights(P(w), Cout, Cin, k, perm_hw, P(wf), P(wd), _s()))


@_op("out")
def conv_fprop(x, N, H, W, Cin, wf, Cout, k, stride, pad, out, bias=None, act=ACT_NONE, slope=0.2, add_src=None):
    _chk(x, wf, out, bias, add_src)
    _lib.check(_L().pcg_conv_fprop(P(x), N, H, W, Cin, P(wf), Cout, k, stride, pad, P(bias), act, _f(slope), P(add_src),
                                   P(out), _s()))


@_op("din")
def conv_dgrad(dout, N, H, W, Cin, wd, Cout, k, stride, pad, din, add_src=None, act_ref=None, ref_act=ACT_NONE,
               ref_slope=0.2):
    """din[N,H,W,Cin] = gradient of conv(geometry N,H,W,Cin -> Cout) wrt its input, given dout."""
    _chk(dout, wd, din, add_src, act_ref)
    _lib.check(_L().pcg_conv_dgrad(P(dout), N, H, W, Cin, P(wd), Cout, k, stride, pad, P(add_src), P(act_ref), ref_act,
                                   _f(ref_slope), P(din), _s()))


def conv_wgrad_scratch_floats(N, H, W, Cin, Cout, k, stride, pad):
    L = _L()
    L.pcg_conv_wgrad_scratch.restype = ctypes.c_longlong
    return int(L.pcg_conv_wgrad_scratch(N, H, W, Cin, Cout, k, stride, pad))


def conv_wgrad_scratch(N, H, W, Cin, Cout, k, stride, pad, device):
    return torch.zeros(conv_wgrad_scratch_floats(N, H, W, Cin, Cout, k, stride, pad), device=device)


@_op("scratch", "dw")
def conv_wgrad(x, dout, N, H, W, Cin, Cout, k, stride, pad, scratch, dw):
    _chk(x, dout, scratch, dw)
    _lib.check(_L().pcg_conv_wgrad(P(x), P(dout), N, H, W, Cin, Cout, k, stride, pad, P(scratch), P(dw), _s()))


def linear_fwd(x, w, out, bias=None, act=ACT_NONE, slope=0.2):
    """out[B,N] = act(x[B,K] @ w[N,K]^T + bias)."""
    B, K = x.shape
    conv_fprop(x, B, 1, 1, K, w, w.shape[0], 1, 1, 0, out, bias, act, slope)


def linear_dgrad(dy, wT, dx, K, act_ref=None, ref_act=ACT_NONE, ref_slope=0.2, add_src=None):
    """dx[B,K] = dy[B,N] @ w[N,K], with wT = w^T [K,N] (pack_weights(..., wd=wT)); optional activation derivative."""
    B, N = dy.shape
    conv_dgrad(dy, B, 1, 1, K, wT, N, 1, 1, 0, dx, add_src, act_ref, ref_act, ref_slope)


def linear_wgrad_small_scratch_floats(M, K, N):
    L = _L()
    L.pcg_linear_wgrad_small_scratch.restype = ctypes.c_longlong
    return int(L.pcg_linear_wgrad_small_scratch(_ll(M), K, N))


@_op("scratch", "dw", "db")
def linear_wgrad_small(x, dy, scratch, dw, db=None):
    """dw[N][K] = dy^T x and db = column sums of dy in one launch (K, N <= 128; pcg_linear_wgrad_small)."""
    M, K = x.shape
    _chk(x, dy, scratch, dw, db)
    _lib.check(_L().pcg_linear_wgrad_small(P(x), P(dy), _ll(M), K, dy.shape[1], P(scratch), P(dw), P(db), _s()))


_SMALL_WGRAD = os.environ.get("PCG_SMALL_WGRAD", "1") != "0"


def linear_wgrad(x, dy, scratch, dw, db=None, stat=None, on_critical_path=False):
    """dw, db of a Linear layer.  Small layers take the one-call kernel (two launches instead of four);
    ``on_critical_path`` forces the primitive operators (kept for A/B measurements, tools/bench_wgrad_small.py)."""
    B, K = x.shape
    N = dy.shape[1]
    if _SMALL_WGRAD and not on_critical_path and N * K <= 4096:      # 64 x 128: 27 us against 14 us for the four launches
        need = linear_wgrad_small_scratch_floats(B, K, N)
        if 0 < need <= scratch.numel() and scratch.data_ptr() % 16 == 0:
            linear_wgrad_small(x, dy, scratch, dw, db)
            return
    conv_wgrad(x, dy, B, 1, 1, K, N, 1, 1, 0, scratch, dw)
    if db is not None:
        colsum(dy, stat, db)


@_op("scratch", "out")
def colsum(a, scratch, out):
    M, C = a.shape[0] if a.dim() == 2 else a.numel() // a.shape[-1], a.shape[-1]
    _chk(a, scratch, out)
    _lib.check(_L().pcg_colsum(P(a), _ll(M), C, P(scratch), P(out), _s()))


# ---------------------------------------------------------------- batch norm
class BNState:
    """Saved statistics of one train-mode BatchNorm forward (C channels)."""

    def __init__(self, C, device):
        z = lambda: torch.zeros(C, device=device)  # noqa: E731
        self.mean, self.rstd, self.scale, self.shift = z(), z(), z(), z()
        self.c12 = torch.zeros(2 * C, device=device)
        self.scratch = stat_scratch(C, device)
        self.scratch2 = stat_scratch(C, device)


@_op("running_mean", "running_var", "nbt", "z", lambda a: [a["st"].mean, a["st"].rstd, a["st"].scale, a["st"].shift, a["st"].scratch])
def bn_train_fwd(y, M, C, gamma, beta, running_mean, running_var, nbt, st, z, act=ACT_NONE, slope=0.2, eps=1e-5,
                 momentum=0.1):
    _chk(y, gamma, beta, z)
    _lib.check(_L().pcg_bn_train_fwd(P(y), _ll(M), C, P(gamma), P(beta), _f(eps), _f(momentum), P(running_mean),
                                     P(running_var), P(nbt), P(st.mean), P(st.rstd), P(st.scale), P(st.shift), act,
                                     _f(slope), P(z), P(st.scratch), _s()))


@_op("dy", "dgamma", "dbeta", "dbias_prev", lambda a: [a["st"].c12, a["st"].scratch, a["st"].scratch2])
def bn_train_bwd(dz, y, M, C, gamma, st, dy, dgamma, dbeta, dbias_prev=None, gscale=1.0, act=ACT_NONE, slope=0.2):
    _chk(dz, y, dy, dgamma, dbeta, dbias_prev)
    _lib.check(_L().pcg_bn_train_bwd(P(dz), P(y), _ll(M), C, P(gamma), P(st.mean), P(st.rstd), P(st.scale), P(st.shift),
                                     _f(gscale), act, _f(slope), P(dy), P(dgamma), P(dbeta), P(dbias_prev), P(st.c12),
                                     P(st.scratch), P(st.scratch2), _s()))


@_op("y", "scale_out")
def bn_eval(x, gamma, beta, rm, rv, y, scale_out=None, eps=1e-5):
    rows, C = x.numel() // x.shape[-1], x.shape[-1]
    _lib.check(_L().pcg_bn_eval(P(x), _ll(rows), C, P(gamma), P(beta), P(rm), P(rv), _f(eps), P(y), P(scale_out), _s()))


@_op("dx")
def scale_cols(dy, scale, dx):
    rows, C = dy.numel() // dy.shape[-1], dy.shape[-1]
    _lib.check(_L().pcg_scale_cols(P(dy), _ll(rows), C, P(scale), P(dx), _s()))


# ---------------------------------------------------------------- elementwise / reductions / losses
@_op("y")
def unary(x, op, y, a=0.0):
    _chk(x, y)
    _lib.check(_L().pcg_unary(P(x), _ll(x.numel()), op, _f(a), P(y), _s()))


@_op("dx")
def unary_bwd(dy, y, op, dx, a=0.0):
    _chk(dy, y, dx)
    _lib.check(_L().pcg_unary_bwd(P(dy), P(y), _ll(y.numel()), op, _f(a), P(dx), _s()))


@_op("out")
def binary(a, b, op, out, alpha=1.0, beta=1.0):
    _chk(a, b, out)
    _lib.check(_L().pcg_binary(P(a), P(b), _ll(a.numel()), op, _f(alpha), _f(beta), P(out), _s()))


@_op("out")
def film_fwd(gamma, n, beta, out, relu=False, res=None):
    """out = [relu](gamma * n + beta) [+ res], one launch."""
    _chk(gamma, n, beta, out, res)
    _lib.check(_L().pcg_film_fwd(P(gamma), P(n), P(beta), P(res), _ll(n.numel()), 1 if relu else 0, P(out), _s()))


@_op("dn", "dgamma", "dbeta")
def film_bwd(df, gamma, n, dn, dgamma, dbeta, accumulate=False):
    """dn = df * gamma ; dgamma (+)= df * n ; dbeta (+)= df, one launch."""
    _chk(df, gamma, n, dn, dgamma, dbeta)
    _lib.check(_L().pcg_film_bwd(P(df), P(gamma), P(n), _ll(n.numel()), 1 if accumulate else 0, P(dn), P(dgamma),
                                 P(dbeta), _s()))


@_op(lambda a: [t for _, t in a["pairs"]])
def transpose_multi(pairs):
    """pairs: list of (W [rows, cols], WT [cols, rows]); one launch per 64 matrices."""
    for i in range(0, len(pairs), 64):
        chunk = pairs[i:i + 64]
        n = len(chunk)
        src = (ctypes.c_void_p * n)(*[w.data_ptr() for w, _ in chunk])
        dst = (ctypes.c_void_p * n)(*[t.data_ptr() for _, t in chunk])
        rows = (ctypes.c_int * n)(*[w.shape[0] for w, _ in chunk])
        cols = (ctypes.c_int * n)(*[w.shape[1] for w, _ in chunk])
        _lib.check(_L().pcg_transpose_multi(n, src, dst, rows, cols, _s()))


@_op(lambda a: [(a["dst"], a["c0_dst"], a["c0_dst"] + a["ncols"])],
     reads=lambda a: [(a["src"], a["c0_src"], a["c0_src"] + a["ncols"])] +
     ([(a["dst"], a["c0_dst"], a["c0_dst"] + a["ncols"])] if a["accumulate"] else []))
def copy_cols(src, c0_src, dst, c0_dst, ncols, alpha=1.0, accumulate=False):
    _chk(src, dst)
    rows = src.shape[0]
    _lib.check(_L().pcg_copy_cols(P(src), src.shape[1], c0_src, P(dst), dst.shape[1], c0_dst, _ll(rows), ncols, _f(alpha),
                                  1 if accumulate else 0, _s()))


@_op("dst")
def onehot(labels, nc, dst, c0=0):
    _chk(labels, dst)
    _lib.check(_L().pcg_onehot(P(labels), _ll(labels.numel()), nc, P(dst), dst.shape[1], c0, _s()))


@_op("out", "dx")
def reduce_scalar(x, out, scale=1.0, absval=False, dx=None, gscale=0.0):
    _chk(x, out, dx)
    _lib.check(_L().pcg_reduce_scalar(P(x), _ll(x.numel()), 1 if absval else 0, _f(scale), P(out), _f(gscale), P(dx), _s()))


@_op("out", "dx")
def rownorm_mean(x, p, out, dx=None, gscale=0.0):
    _chk(x, out, dx)
    _lib.check(_L().pcg_rownorm_mean(P(x), _ll(x.shape[0]), x.shape[1], p, P(out), _f(gscale), P(dx), _s()))


@_op("out_loss", "dz", "out_aux")
def gan_loss(z, kind, t, out_loss, dz, wgt=1.0, out_aux=None):
    _chk(z, out_loss, dz, out_aux)
    _lib.check(_L().pcg_gan_loss(P(z), z.numel(), kind, _f(t), _f(wgt), P(out_loss), P(out_aux), P(dz), _s()))


@_op("out")
def combine(terms, out):
    """out[0] = sum coeff * scalar_tensor[0] over up to 6 (coeff, tensor) pairs."""
    n = len(terms)
    coeffs = (ctypes.c_float * n)(*[float(c) for c, _ in terms])
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for _, t in terms])
    _lib.check(_L().pcg_combine_scalars(n, coeffs, ptrs, P(out), _s()))


@_op(lambda a: [a["G"].data, a["G"].grad, a["G"].m, a["G"].v, a["G"].step, a["D"].data, a["D"].grad, a["D"].m, a["D"].v, a["D"].step, a["scal"]])
def mlp_gan_step(B, z_dim, label_dim, hidden, real, real_oh, z1, oh1, z2, oh2, G, D, lr, scal):
    """One whole iteration of the two-layer MLP GAN in one cluster launch (csrc/mlp_gan.cu); G, D are FlatParams."""
    _chk(real, real_oh, z1, oh1, z2, oh2, G.data, G.grad, G.m, G.v, D.data, D.grad, D.m, D.v, scal)
    _lib.check(_L().pcg_mlp_gan_step(B, z_dim, label_dim, hidden, P(real), P(real_oh), P(z1), P(oh1), P(z2), P(oh2),
                                     P(G.data), P(G.grad), P(G.m), P(G.v), P(G.step), P(D.data), P(D.grad), P(D.m), P(D.v),
                                     P(D.step), _f(lr), P(scal), _s()))


@_op("u", "v", "Wn", "sigma", "WnT", "us", "vs")
def spectral_norm_fwd(W, u, v, Wn, sigma, do_iter=True, eps=1e-12, WnT=None, us=None, vs=None):
    """One power iteration (in place on u, v), sigma, Wn = W / sigma; optionally Wn^T and snapshots of u, v."""
    N, K = W.shape
    _chk(W, u, v, Wn, sigma, WnT, us, vs)
    _lib.check(_L().pcg_spectral_norm_fwd2(P(W), N, K, P(u), P(v), _f(eps), 1 if do_iter else 0, P(Wn), P(WnT), P(us),
                                           P(vs), P(sigma), _s()))


@_op("dW")
def spectral_norm_bwd(dWn, Wn, u, v, sigma, dW):
    N, K = Wn.shape
    _lib.check(_L().pcg_spectral_norm_bwd(P(dWn), P(Wn), N, K, P(u), P(v), P(sigma), P(dW), _s()))


@_op("y")
def gumbel_softmax_fwd(logits, g, tau, y):
    rows, n = logits.shape
    _lib.check(_L().pcg_gumbel_softmax_fwd(P(logits), P(g), _ll(rows), n, _f(tau), P(y), _s()))


@_op("y")
def onehot_argmax(x, y):
    rows, n = x.shape
    _chk(x, y)
    _lib.check(_L().pcg_onehot_argmax(P(x), _ll(rows), n, P(y), _s()))


@_op("dl")
def softmax_bwd(dy, y, tau, dl):
    rows, n = y.shape
    _lib.check(_L().pcg_softmax_bwd(P(dy), P(y), _ll(rows), n, _f(tau), P(dl), _s()))


@_op("loss", "dlogits")
def ce_loss(logits, target, loss, dlogits, wgt=1.0):
    B, NC = logits.shape
    _lib.check(_L().pcg_ce_loss(P(logits), P(target), B, NC, _f(wgt), P(loss), P(dlogits), _s()))


@_op("y", "mean", "rstd")
def instnorm_fwd(x, N, HW, C, gamma, beta, y, mean, rstd, act=ACT_NONE, slope=0.2, eps=1e-5):
    _chk(x, gamma, beta, y, mean, rstd)
    _lib.check(_L().pcg_instnorm_fwd(P(x), N, HW, C, P(gamma), P(beta), _f(eps), act, _f(slope), P(y), P(mean), P(rstd),
                                     _s()))


@_op("dx", "dgamma_part", "dbeta_part")
def instnorm_bwd(gy, x, mean, rstd, gamma, N, HW, C, dx, act_ref=None, act=ACT_NONE, slope=0.2, add_src=None,
                 dgamma_part=None, dbeta_part=None):
    _chk(gy, x, mean, rstd, gamma, dx, act_ref, add_src, dgamma_part, dbeta_part)
    _lib.check(_L().pcg_instnorm_bwd(P(gy), P(act_ref), act, _f(slope), P(x), P(mean), P(rstd), P(gamma), N, HW, C,
                                     P(add_src), P(dx), P(dgamma_part), P(dbeta_part), _s()))


@_op("gy_bar", "x_bar", "dgamma_part")
def instnorm_bwd_bwd(q, gy, x, mean, rstd, gamma, N, HW, C, gy_bar, x_bar, act_ref=None, act=ACT_NONE, slope=0.2,
                     dgamma_part=None):
    _chk(q, gy, x, mean, rstd, gamma, gy_bar, x_bar, act_ref, dgamma_part)
    _lib.check(_L().pcg_instnorm_bwd_bwd(P(q), P(gy), P(act_ref), act, _f(slope), P(x), P(mean), P(rstd), P(gamma), N,
                                         HW, C, P(gy_bar), P(x_bar), P(dgamma_part), _s()))


@_op(lambda a: [a["dst"] if a.get("inverse") else (a["dst"], a["c0"], a["c0"] + a["R"] * a["C"])])
def flatten_nchw(src, B, R, C, dst, ld, c0=0, inverse=False):
    _chk(src, dst)
    _lib.check(_L().pcg_flatten_nchw(P(src), B, R, C, P(dst), ld, c0, 1 if inverse else 0, _s()))


def frozen_mlp_parts(dims, B):
    """Floats of the per-CTA loss parts of ``frozen_mlp_ce_grad``; -1 when the shape is not supported."""
    arr = (ctypes.c_int * len(dims))(*dims)
    return int(_L().pcg_frozen_mlp_parts(len(dims) - 1, arr, B))


@_op("logits", "loss_part", "dx")
def frozen_mlp_ce_grad(weights, weights_t, biases, x, target, loss_part, dx, wgt=1.0, slope=0.0, logits=None, mean_output=False):
    """Frozen MLP classifier: forward, cross-entropy against ``target`` and the gradient with respect to ``x`` in one launch
    (pcg_frozen_mlp_ce_grad).  weights[j] is the torch [out][in] matrix of layer j, weights_t[j] its transpose;
    LeakyReLU(slope) between layers.  mean_output: the loss is the mean of the network's outputs (a critic score) instead
    of the cross-entropy, ``target`` is not used."""
    B, d0 = x.shape
    dims = [d0] + [w.shape[0] for w in weights]
    _chk(x, loss_part, dx, logits, *weights, *weights_t, *biases)
    L = len(weights)
    ptrs = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])  # noqa: E731
    _lib.check(_L().pcg_frozen_mlp_ce_grad(L, (ctypes.c_int * (L + 1))(*dims), ptrs(weights), ptrs(weights_t), ptrs(biases),
                                           _f(slope), P(x), P(target), 1 if mean_output else 0, B, _f(wgt), P(logits),
                                           P(loss_part), P(dx), _s()))


@_op("logits", "loss_part", "dx", "acts", "grads")
def mlp_fwd_bwd(weights, weights_t, biases, x, target, loss_part, dx, acts, grads, wgt=1.0, slope=0.0, logits=None,
                mean_output=False):
    """``frozen_mlp_ce_grad`` that also stores the hidden activations (``acts[j]`` [B, width]) and the gradients of their
    pre-activations (``grads[j]``) for the caller's weight gradients (pcg_mlp_fwd_bwd)."""
    B, d0 = x.shape
    dims = [d0] + [w.shape[0] for w in weights]
    _chk(x, loss_part, dx, logits, *weights, *weights_t, *biases, *acts, *grads)
    L = len(weights)
    ptrs = lambda ts, n: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
    _lib.check(_L().pcg_mlp_fwd_bwd(L, (ctypes.c_int * (L + 1))(*dims), ptrs(weights, L), ptrs(weights_t, L), ptrs(biases, L),
                                    _f(slope), P(x), P(target), 1 if mean_output else 0, B, _f(wgt), P(logits), P(loss_part),
                                    P(dx), ptrs(acts, L - 1), ptrs(grads, L - 1), _s()))


def film_layer_supported(M, H):
    return bool(_L().pcg_film_layer_supported(_ll(M), H))


@_op("running_mean", "running_var", "nbt", "u", "n", "out",
     lambda a: [a["st"].mean, a["st"].rstd, a["st"].scale, a["st"].shift, a["st"].scratch])
def film_layer_fwd(x, W, bias, gamma, beta, running_mean, running_var, nbt, st, fg, fb, u, n, out, res=None, relu=False,
                   eps=1e-5, momentum=0.1):
    """One call: Linear(H, H) -> BatchNorm1d (training) -> FiLM -> ReLU (relu=True) or ``res +`` (pcg_film_layer_fwd)."""
    M, H = x.shape
    _chk(x, W, bias, gamma, beta, running_mean, running_var, fg, fb, u, n, out, res)
    _lib.check(_L().pcg_film_layer_fwd(P(x), _ll(M), H, P(W), P(bias), P(gamma), P(beta), _f(eps), _f(momentum),
                                       P(running_mean), P(running_var), P(nbt), P(st.mean), P(st.rstd), P(st.scale),
                                       P(st.shift), P(fg), P(fb), P(res), 1 if relu else 0, P(u), P(n), P(out),
                                       P(st.scratch), _s()))


class _FilmHalf(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("W", "bias", "gamma", "beta", "running_mean", "running_var", "nbt", "mean", "rstd",
                                               "scale", "shift", "fg", "fb", "res", "u", "n", "out", "scratch")]


def _film_chain_writes(a):
    w = []
    for (W, bias, gamma, beta, rm, rv, nbt, st, fg, fb, res, u, n, out) in a["halves"]:
        w += [rm, rv, nbt, st.mean, st.rstd, st.scale, st.shift, st.scratch, u, n, out]
    return w


@_op(_film_chain_writes)
def film_chain_fwd(x, halves, eps=1e-5, momentum=0.1):
    """``halves``: list of (W, bias, gamma, beta, running_mean, running_var, nbt, st, fg, fb, res_or_None, u, n, out), each
    half block fed by the previous one's ``out``; n + 1 launches instead of 2 n (pcg_film_chain_fwd)."""
    M, H = x.shape
    arr = (_FilmHalf * len(halves))()
    for k, (W, bias, gamma, beta, rm, rv, nbt, st, fg, fb, res, u, n, out) in enumerate(halves):
        _chk(W, bias, gamma, beta, rm, rv, fg, fb, res, u, n, out)
        for name, t in (("W", W), ("bias", bias), ("gamma", gamma), ("beta", beta), ("running_mean", rm), ("running_var", rv),
                        ("nbt", nbt), ("mean", st.mean), ("rstd", st.rstd), ("scale", st.scale), ("shift", st.shift), ("fg", fg),
                        ("fb", fb), ("res", res), ("u", u), ("n", n), ("out", out), ("scratch", st.scratch)):
            setattr(arr[k], name, None if t is None else t.data_ptr())
    _lib.check(_L().pcg_film_chain_fwd(P(x), _ll(M), H, len(halves), arr, _f(eps), _f(momentum), _s()))


@_op("dfg", "dfb", "du", "dx", "dgamma", "dbeta", lambda a: [a["st"].scratch2])
def film_layer_bwd(d_f, fg, n, u, st, gamma, W, dfg, dfb, du, dx, dgamma, dbeta, add_src=None, act_ref=None,
                   accumulate=False):
    M, H = d_f.shape
    _chk(d_f, fg, n, u, gamma, W, dfg, dfb, du, dx, dgamma, dbeta, add_src, act_ref)
    _lib.check(_L().pcg_film_layer_bwd(P(d_f), _ll(M), H, P(fg), P(n), P(u), P(st.mean), P(st.rstd), P(gamma), P(W),
                                       P(add_src), P(act_ref), 1 if accumulate else 0, P(dfg), P(dfb), P(du), P(dx),
                                       P(dgamma), P(dbeta), P(st.scratch2), _s()))


@_op("y")
def bias_act(x, C, bias, y, tanh_out=False):
    _chk(x, bias, y)
    _lib.check(_L().pcg_bias_act(P(x), _ll(x.numel() // C), C, P(bias), 1 if tanh_out else 0, P(y), _s()))


@_op("dst")
def dilate(src, N, Ho, Wo, C, stride, off, Hp, Wp, dst):
    _chk(src, dst)
    _lib.check(_L().pcg_dilate(P(src), N, Ho, Wo, C, stride, off, Hp, Wp, P(dst), _s()))


@_op("wc")
def pack_dgrad_classes(w, k, wc):
    """torch OIHW weight -> wc [4][Cin][2*2][Cout], the forward weights of the four parity-class convolutions whose
    interleaved results are the stride-2 data gradient (include/pcg.h)."""
    _chk(w, wc)
    _lib.check(_L().pcg_pack_dgrad_classes(P(w), w.shape[0], w.shape[1], k, P(wc), _s()))


@_op("dx")
def parity_interleave(src, N, Hc, Wc, C, pad, H, W, dx, stacked=False):
    _chk(src, dx)
    _lib.check(_L().pcg_parity_interleave(P(src), N, Hc, Wc, C, pad, H, W, 1 if stacked else 0, P(dx), _s()))


@_op("out", "gbar", "norms")
def gp_penalty(g, B, D, lam, out, gbar, norms=None):
    _chk(g, out, gbar, norms)
    _lib.check(_L().pcg_gp_penalty(P(g), B, D, _f(lam), P(out), P(gbar), P(norms), _s()))


def cf_scratch(device):
    return torch.zeros(int(_L().pcg_cf_scratch_floats()), device=device)


@_op("x_cf", "scratch")
def cf_apply(x, residual, x_cf, scratch, lo=-1.0, hi=1.0):
    _chk(x, residual, x_cf, scratch)
    _lib.check(_L().pcg_cf_apply(P(x), P(residual), _ll(x.numel()), _f(lo), _f(hi), P(x_cf), P(scratch), _s()))


@_op("out3")
def cf_metrics(logits, y_true, y_target, scratch, n_elems, out3):
    B, NC = logits.shape
    _chk(logits, y_true, y_target, scratch, out3)
    _lib.check(_L().pcg_cf_metrics(P(logits), P(y_true), P(y_target), B, NC, P(scratch), _ll(n_elems), P(out3), _s()))


@_op("mask", "rng_state")
def dropout_mask(mask, p, channelwise=False, seed=0, rng_state=None):
    """mask [rows, ..., C] = Bernoulli(1 - p) / (1 - p); channelwise: one draw per (row, channel) (nn.Dropout2d on NHWC)."""
    _chk(mask, rng_state)
    rows, C = mask.shape[0], mask.shape[-1]
    inner = mask.numel() // (rows * C)
    _lib.check(_L().pcg_dropout_mask(_ll(rows), inner, C, _f(p), 1 if channelwise else 0,
                                     ctypes.c_ulonglong(seed & (2 ** 64 - 1)), P(rng_state), P(mask), _s()))


@_op("mask", "target", "rng_state")
def build_mask(B, C, H, W, patch, num_modifiable_patches, mask, target=None, num_classes=10, seed=0, rng_state=None):
    """One launch: random patch mask [B,C,H,W] (+ target draw [B] int64); see include/pcg.h pcg_build_mask."""
    _chk(mask, target, rng_state)
    k = -1 if num_modifiable_patches is None else int(num_modifiable_patches)
    _lib.check(_L().pcg_build_mask(B, C, H, W, patch, k, num_classes, ctypes.c_ulonglong(seed & (2 ** 64 - 1)),
                                   P(rng_state), P(mask), P(target), _s()))


@_op("loss", "dlogits", "correct")
def ce_loss_weighted(logits, target, class_weights, loss, dlogits=None, correct=None):
    B, NC = logits.shape
    _chk(logits, target, class_weights, loss, dlogits, correct)
    _lib.check(_L().pcg_ce_loss_weighted(P(logits), P(target), P(class_weights), B, NC, P(loss), P(dlogits), P(correct), _s()))


@_op("p", "m", "v", "step")
def adamw(p, g, m, v, step, lr_dev, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    _chk(p, g, m, v, step, lr_dev)
    _lib.check(_L().pcg_adamw_flat(P(p), P(g), P(m), P(v), _ll(p.numel()), P(step), P(lr_dev), _f(beta1), _f(beta2), _f(eps),
                                   _f(weight_decay), _s()))


@_op("p", "m", "v", "step")
def adam(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    _chk(p, g, m, v, step)
    _lib.check(_L().pcg_adam_flat(P(p), P(g), P(m), P(v), _ll(p.numel()), P(step), _f(lr), _f(beta1), _f(beta2), _f(eps),
                                  _f(grad_scale), _s()))


class FlatParams:
    """Parameters of a small network as views into one flat fp32 buffer (+ grads, Adam state), so that the
    optimizer is one kernel and ``nn.Parameter`` tensors can alias the slices (state_dict keeps working)."""

    def __init__(self, named_shapes, device):
        self.names = [n for n, _ in named_shapes]
        self.shapes = {n: tuple(s) for n, s in named_shapes}
        offs, off = {}, 0
        for n, s in named_shapes:
            offs[n] = off
            numel = 1
            for d in s:
                numel *= d
            off += (numel + 3) // 4 * 4
        self.offsets, self.size = offs, off
        self.data = torch.zeros(off, device=device)
        self.grad = torch.zeros(off, device=device)
        self.m = torch.zeros(off, device=device)
        self.v = torch.zeros(off, device=device)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)

    def _view(self, buf, n):
        s = self.shapes[n]
        numel = 1
        for d in s:
            numel *= d
        return buf[self.offsets[n]:self.offsets[n] + numel].view(s)

    def p(self, n):
        return self._view(self.data, n)

    def g(self, n):
        return self._view(self.grad, n)

    def load(self, tensors):
        for n, t in tensors.items():
            self.p(n).copy_(t.to(self.data.device, torch.float32))
        return self

    def adopt(self, module):
        """Re-points the module's parameters at the flat buffer (values copied in first)."""
        for n, prm in module.named_parameters():
            v = self.p(n)
            v.copy_(prm.detach().to(v.device, torch.float32))
            prm.data = v
            if prm.requires_grad:
                prm.grad = self.g(n)
        return self

    def aliases(self, module):
        """True while every parameter of ``module`` still points into this arena (a later plan may have re-adopted
        the module into its own arena: a forward-only plan cached on the module must not be reused then)."""
        try:
            return all(prm.data_ptr() == self.p(n).data_ptr() for n, prm in module.named_parameters())
        except KeyError:
            return False

    def adam_step(self, lr, grad_scale=1.0):
        adam(self.data, self.grad, self.m, self.v, self.step, lr, grad_scale=grad_scale)
