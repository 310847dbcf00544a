"""Building blocks of the tabular step plans: dense layers (optionally spectrally normalised), train-mode
BatchNorm1d, the spectral-norm MLP critic shared by both tabular CounteRGANs, graph capture.  Every method only
enqueues libpcg kernels (pcg_b200.ops); gradients are written into FlatParams arenas."""
import torch

from .. import dataflow
from .. import graphs
from .. import ops as K


class Ctx:
    """Shared scratch of one plan."""

    def __init__(self, B, dev, max_dim=256):
        self.B, self.dev = B, dev
        self.stat = K.stat_scratch(max(max_dim, 4), dev)
        self._wsc_need, self._wsc = K.conv_wgrad_scratch_floats(B, 1, 1, max_dim, max_dim, 1, 1, 0), None

    def reserve_wgrad(self, k_in, n_out):
        """Every layer registers its wgrad geometry: the split-K partial buffer depends on the shape (the Cout==1 path
        keeps one partial row per block), so the shared scratch is sized for the largest request."""
        assert self._wsc is None, "layers must be created before the first wgrad"
        self._wsc_need = max(self._wsc_need, K.conv_wgrad_scratch_floats(self.B, 1, 1, k_in, n_out, 1, 1, 0))

    @property
    def wsc(self):
        if self._wsc is None:
            self._wsc = torch.zeros(self._wsc_need + 1024, device=self.dev)
        return self._wsc

    def z(self, *s):
        return torch.zeros(*s, device=self.dev)


class _OwnScratch:
    """Per-layer weight-gradient / column-sum scratch: layers that share one buffer are serialised by it (write after
    write), and the weight gradients are exactly the launches the data-flow capture can take off the critical path."""

    def _own_scratch(self, ctx, k_in, n_out):
        need = max(K.conv_wgrad_scratch_floats(ctx.B, 1, 1, k_in, n_out, 1, 1, 0),
                   K.linear_wgrad_small_scratch_floats(ctx.B, k_in, n_out))       # -1 when the one-launch kernel does not apply
        self.wsc = torch.zeros(need + 1024, device=ctx.dev)
        self.stat = K.stat_scratch(max(n_out, 4), ctx.dev)


class Dense(_OwnScratch):
    def __init__(self, ctx, flat, name, k_in, n_out):
        self.ctx, self.flat, self.name, self.k, self.n = ctx, flat, name, k_in, n_out
        self.wT = ctx.z(k_in, n_out)
        ctx.reserve_wgrad(k_in, n_out)
        self._own_scratch(ctx, k_in, n_out)

    def W(self):
        return self.flat.p(self.name + ".weight")

    def b(self):
        return self.flat.p(self.name + ".bias")

    def refresh(self):
        K.pack_weights(self.W(), 1, wd=self.wT)

    def fwd(self, x, out, act=K.ACT_NONE, slope=0.2):
        K.linear_fwd(x, self.W(), out, self.b(), act, slope)

    def dgrad(self, dy, dx, **kw):
        K.linear_dgrad(dy, self.wT, dx, self.k, **kw)

    def wgrad(self, x, dy):
        K.linear_wgrad(x, dy, self.wsc, self.flat.g(self.name + ".weight"), self.flat.g(self.name + ".bias"), self.stat)


class BN1d:
    """Train-mode BatchNorm1d over the batch rows; running buffers are caller tensors (module buffers)."""

    def __init__(self, ctx, flat, name, C):
        self.ctx, self.flat, self.name, self.C = ctx, flat, name, C
        self.rm, self.rv = ctx.z(C), torch.ones(C, device=ctx.dev)
        self.nbt = torch.zeros((), dtype=torch.int64, device=ctx.dev)
        self.st = K.BNState(C, ctx.dev)

    def fwd(self, y, z, act=K.ACT_NONE, training=True):
        g, b = self.flat.p(self.name + ".weight"), self.flat.p(self.name + ".bias")
        if training:
            K.bn_train_fwd(y, y.shape[0], self.C, g, b, self.rm, self.rv, self.nbt, self.st, z, act=act)
        else:
            K.bn_eval(y, g, b, self.rm, self.rv, z)
            if act == K.ACT_RELU:
                K.unary(z, K.RELU, z)

    def bwd(self, dz, y, dy, act=K.ACT_NONE):
        K.bn_train_bwd(dz, y, y.shape[0], self.C, self.flat.p(self.name + ".weight"), self.st, dy,
                       self.flat.g(self.name + ".weight"), self.flat.g(self.name + ".bias"), act=act)


class SNDense(_OwnScratch):
    """spectral_norm(nn.Linear) (torch/nn/utils/spectral_norm.py): one power iteration per forward call in train mode,
    in place on the module's u / v; each of the plan's forward passes keeps its own snapshot (u, v, sigma, W/sigma)
    because torch's graph holds clones taken at call time."""

    def __init__(self, ctx, flat, name, k_in, n_out, passes=2):
        self.ctx, self.flat, self.name, self.k, self.n = ctx, flat, name, k_in, n_out
        z = ctx.z
        ctx.reserve_wgrad(k_in, n_out)
        self._own_scratch(ctx, k_in, n_out)
        self.u, self.v = z(n_out), z(k_in)
        self.Wn = [z(n_out, k_in) for _ in range(passes)]
        self.WnT = [z(k_in, n_out) for _ in range(passes)]
        self.sigma = [z(1) for _ in range(passes)]
        self.us = [z(n_out) for _ in range(passes)]
        self.vs = [z(k_in) for _ in range(passes)]
        self.dWn = [z(n_out, k_in) for _ in range(passes)]

    def W(self):
        return self.flat.p(self.name + ".weight_orig")

    def b(self):
        return self.flat.p(self.name + ".bias")

    def normalise(self, p, iterate=True):
        K.spectral_norm_fwd(self.W(), self.u, self.v, self.Wn[p], self.sigma[p], do_iter=iterate, WnT=self.WnT[p],
                            us=self.us[p], vs=self.vs[p])

    def fwd(self, x, out, p, act=K.ACT_NONE, slope=0.2):
        K.linear_fwd(x, self.Wn[p], out, self.b(), act, slope)

    def dgrad(self, dy, dx, p, **kw):
        K.linear_dgrad(dy, self.WnT[p], dx, self.k, **kw)

    def wgrad(self, x, dy, p, garena):
        """garena(name) -> gradient slot.  dW_orig = (dWn - <dWn, Wn> u v^T) / sigma."""
        K.linear_wgrad(x, dy, self.wsc, self.dWn[p], garena(self.name + ".bias"), self.stat)
        K.spectral_norm_bwd(self.dWn[p], self.Wn[p], self.us[p], self.vs[p], self.sigma[p], garena(self.name + ".weight_orig"))


class Critic:
    """cat[x, onehot] -> 4 spectral-norm Linear layers with LeakyReLU(0.2) between
    (moons/models/discriminator.py:6-22, house_sales_kc_usa/models/discriminator.py:5-20)."""

    def __init__(self, ctx, dims, device):
        self.ctx, self.dims = ctx, dims
        B = ctx.B
        names = []
        for i, (a, b) in enumerate(dims):
            names += [(f"net.{2 * i}.bias", (b,)), (f"net.{2 * i}.weight_orig", (b, a))]
        self.flat = K.FlatParams(names, device)
        self.grad2 = torch.zeros_like(self.flat.grad)
        self.layers = [SNDense(ctx, self.flat, f"net.{2 * i}", a, b) for i, (a, b) in enumerate(dims)]
        self.din = [ctx.z(B, dims[0][0]) for _ in range(2)]
        self.h = [[ctx.z(B, b) for (_, b) in dims] for _ in range(2)]
        self.dh = [[ctx.z(B, b) for (_, b) in dims] for _ in range(2)]
        self.ddin = ctx.z(B, dims[0][0])
        self.ddin_scratch = [ctx.z(B, dims[0][0]) for _ in range(2)]

    def adopt(self, module):
        self.flat.adopt(module)
        for i, L in enumerate(self.layers):
            lin = module.net[2 * i]
            L.u.copy_(lin.weight_u)
            L.v.copy_(lin.weight_v)
            lin._buffers["weight_u"], lin._buffers["weight_v"] = L.u, L.v

    def fwd(self, x, onehot, p, iterate=True):
        """Forward pass p (0 or 1): power iteration + normalisation of every layer, then the MLP; returns [B,1]."""
        xd = x.shape[1]
        K.copy_cols(x, 0, self.din[p], 0, xd)
        K.copy_cols(onehot, 0, self.din[p], xd, onehot.shape[1])
        h = self.din[p]
        for i, L in enumerate(self.layers):
            L.normalise(p, iterate)
            L.fwd(h, self.h[p][i], p, K.ACT_LRELU if i < 3 else K.ACT_NONE, 0.2)
            h = self.h[p][i]
        return h

    def bwd(self, dz, p, garena=None, want_dx=False):
        """dz: gradient wrt the critic output [B,1].  Weight gradients go to garena (None = data gradient only)."""
        d = dz
        for i in range(3, -1, -1):
            L = self.layers[i]
            xin = self.din[p] if i == 0 else self.h[p][i - 1]
            if garena is not None:
                L.wgrad(xin, d, p, garena)
            if i == 0:
                if want_dx:
                    L.dgrad(d, self.ddin, p)
                break
            L.dgrad(d, self.dh[p][i - 1], p, act_ref=self.h[p][i - 1], ref_act=K.ACT_LRELU, ref_slope=0.2)
            d = self.dh[p][i - 1]
        return self.ddin

    def input_gradient(self, x, onehot, p, wgt, loss_part):
        """The generator's adversarial term in ONE launch (csrc/frozen_mlp.cu): forward pass p with the critic's current
        weights (power iteration included, as ``fwd``), the loss ``mean(score)`` and d (wgt * loss) / d [x, onehot] -> ddin.
        The scores land in h[p][-1] like ``fwd``'s.  Only when ``fused_ok``."""
        xd = x.shape[1]
        K.copy_cols(x, 0, self.din[p], 0, xd)
        K.copy_cols(onehot, 0, self.din[p], xd, onehot.shape[1])
        for L in self.layers:
            L.normalise(p, True)
        K.frozen_mlp_ce_grad([L.Wn[p] for L in self.layers], [L.WnT[p] for L in self.layers], [L.b() for L in self.layers],
                             self.din[p], None, loss_part, self.ddin, wgt=wgt, slope=0.2, logits=self.h[p][-1],
                             mean_output=True)
        return self.h[p][-1], self.ddin

    def update_pass(self, x, onehot, p, sign, garena, loss_part, dz):
        """One pass of the critic's OWN update (forward, loss sign * mean(score), every weight gradient into ``garena``):
        the forward / backward chain as one launch (pcg_mlp_fwd_bwd stores the activations and pre-activation gradients),
        then the per-layer weight gradients side by side.  dz: the constant gradient of the scores, sign / B, [B,1]."""
        xd = x.shape[1]
        K.copy_cols(x, 0, self.din[p], 0, xd)
        K.copy_cols(onehot, 0, self.din[p], xd, onehot.shape[1])
        for L in self.layers:
            L.normalise(p, True)
        n = len(self.layers)
        K.mlp_fwd_bwd([L.Wn[p] for L in self.layers], [L.WnT[p] for L in self.layers], [L.b() for L in self.layers],
                      self.din[p], None, loss_part, self.ddin_scratch[p], self.h[p][:n - 1], self.dh[p][:n - 1], wgt=sign,
                      slope=0.2, logits=self.h[p][-1], mean_output=True)
        for i in range(n - 1, -1, -1):
            xin = self.din[p] if i == 0 else self.h[p][i - 1]
            self.layers[i].wgrad(xin, dz if i == n - 1 else self.dh[p][i], p, garena)
        return self.h[p][-1]

    @property
    def fused_ok(self):
        return K.frozen_mlp_parts([self.dims[0][0]] + [b for _, b in self.dims], self.ctx.B) > 0

    def g1(self, n):
        return self.flat.g(n)

    def g2(self, n):
        return self.flat._view(self.grad2, n)


class GraphStep:
    """Runs ``body`` eagerly once on a state snapshot (module loads, attribute setting), then replays a CUDA graph.
    ``parallel`` (default, PCG_DATAFLOW=0 turns it off): the graph is captured from the body's recorded data flow
    (pcg_b200.dataflow) - every operator waits only for the operators whose memory it touches - instead of as one chain."""

    def __init__(self, body, state_tensors, refresh, use_graph=True, parallel=None):
        import os
        self.body, self.state, self.refresh, self.use_graph, self.graph = body, state_tensors, refresh, use_graph, None
        self.parallel = (os.environ.get("PCG_DATAFLOW", "1") != "0") if parallel is None else bool(parallel)
        self.program = None

    def __call__(self):
        if not self.use_graph:
            self.body()
            return
        if self.graph is None:
            snap = [t.clone() for t in self.state()]
            self.body()
            torch.cuda.synchronize()
            for dst, src in zip(self.state(), snap):
                dst.copy_(src)
            self.refresh()
            torch.cuda.synchronize()
            if self.parallel:
                self.program = dataflow.record(self.body)
                self.graph = graphs.capture(self.program.emit)
            else:
                self.graph = graphs.capture(self.body)
        self.graph.replay()
