"""Native mirror of the two moons MLP GANs.

    conditional_gan/moons/make_moons_cgan.py   Generator(z_dim, label_dim, hidden_dim).forward(z, label_onehot)   :36-46
                                               Discriminator(label_dim, hidden_dim).forward(x, label_onehot)     :49-60
                                               training loop (script body)                                        :83-135
    simple_gan/moons/make_moons_gan.py         build_generator / build_discriminator / train_gan                  :33-93

One iteration = generator forward, discriminator forward on real+fake (batched), saturating log losses, both
backward passes and both Adam updates.  Default: ONE launch of ``pcg_mlp_gan_step`` (csrc/mlp_gan.cu: a thread-block
cluster with all weights in shared memory, warp-shuffle reductions, gradients summed through distributed shared
memory).  Shapes outside that kernel's envelope (hidden != 128, batch > 1024, ...) and ``fused=False`` use the same
iteration composed from libpcg's primitive operators and replayed as one CUDA graph.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import graphs
from .. import ops as K


class _Net(nn.Module):
    """Linear-ReLU-Linear container with the reference's parameter names; forward runs in libpcg."""

    def __init__(self, in_dim, hidden, out_dim, sigmoid, prefix_net=True):
        super().__init__()
        layers = [nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, out_dim)]
        if sigmoid:
            layers.append(nn.Sigmoid())
        self.net = nn.Sequential(*layers)
        self._sigmoid = sigmoid

    def _run(self, x):
        if not x.is_cuda:
            raise RuntimeError("pcg_b200: inputs must be CUDA tensors (there is no CPU fallback)")
        x = x.detach().float().contiguous()
        w1, b1, w2, b2 = (p.detach().contiguous() for p in self.net.parameters())
        h = torch.empty(x.shape[0], w1.shape[0], device=x.device)
        y = torch.empty(x.shape[0], w2.shape[0], device=x.device)
        K.linear_fwd(x, w1, h, b1, K.ACT_RELU)
        K.linear_fwd(h, w2, y, b2)
        if self._sigmoid:
            K.unary(y, K.SIGMOID, y)
        return y


class Generator(_Net):
    def __init__(self, z_dim, label_dim, hidden_dim):
        super().__init__(z_dim + label_dim, hidden_dim, 2, sigmoid=False)

    def forward(self, z, label_onehot=None):
        return self._run(z if label_onehot is None else torch.cat([z, label_onehot], 1))


class Discriminator(_Net):
    def __init__(self, label_dim, hidden_dim):
        super().__init__(2 + label_dim, hidden_dim, 1, sigmoid=True)

    def forward(self, x, label_onehot=None):
        return self._run(x if label_onehot is None else torch.cat([x, label_onehot], 1))


def build_generator(z_dim, hidden_dim, device="cuda"):
    """make_moons_gan.py:33-38 (an nn.Sequential there; the mirror keeps ``net.`` in front of the same keys)."""
    return Generator(z_dim, 0, hidden_dim).to(device)


def build_discriminator(hidden_dim, device="cuda"):
    return Discriminator(0, hidden_dim).to(device)


class MlpGanPlan:
    """Fixed-batch native plan of one G+D iteration; ``label_dim = 0`` gives the unconditional GAN."""

    def __init__(self, batch, z_dim, label_dim, hidden, device, lr=1e-3, use_graph=True, fused=None):
        self.B, self.zd, self.ld, self.H, self.lr = batch, z_dim, label_dim, hidden, lr
        can_fuse = hidden == 128 and label_dim <= 2 and z_dim + label_dim <= 36 and batch <= 1024
        if fused and not can_fuse:
            raise ValueError("pcg_mlp_gan_step needs hidden == 128, label_dim <= 2, z_dim + label_dim <= 36, batch <= 1024")
        self.fused = can_fuse if fused is None else fused
        self._fn = None
        dev = self.dev = torch.device(device)
        gi, di = z_dim + label_dim, 2 + label_dim
        self.gi, self.di = gi, di
        names = lambda i, o: [("net.0.weight", (hidden, i)), ("net.0.bias", (hidden,)), ("net.2.weight", (o, hidden)),  # noqa: E731
                              ("net.2.bias", (o,))]
        self.G, self.D = K.FlatParams(names(gi, 2), dev), K.FlatParams(names(di, 1), dev)
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        B, H = batch, hidden
        # transposed weights for the data gradients
        self.GW2t, self.DW1t, self.DW2t = z(H, 2), z(di, H), z(H, 1)
        # static inputs
        self.real, self.z1, self.z2 = z(B, 2), z(B, z_dim), z(B, z_dim)
        self.real_oh, self.oh1, self.oh2 = (z(B, label_dim) if label_dim else None for _ in range(3))
        # activations
        self.gin, self.gh, self.fake = z(B, gi), z(B, H), z(B, 2)
        self.din, self.dh, self.zl, self.dz = z(2 * B, di), z(2 * B, H), z(2 * B, 1), z(2 * B, 1)
        self.ddh, self.ddin, self.dfake, self.dgh = z(2 * B, H), z(B, di), z(B, 2), z(B, H)
        self.scal = z(8)            # 0 loss_D, 1 loss_G, 2 loss_D real term, 3 loss_D fake term, 4/5 mean D(real)/D(fake)
        self.stat = K.stat_scratch(max(H, 4), dev)
        self.wsc = torch.zeros(max(int(K.conv_wgrad_scratch(2 * B, 1, 1, a, b, 1, 1, 0, dev).numel())
                                   for a, b in ((gi, H), (H, 2), (di, H), (H, 1))), device=dev)
        self.use_graph, self.graph = use_graph, None
        self.refresh()

    def refresh(self):
        """Re-derive the transposed weights after the parameters changed."""
        K.pack_weights(self.G.p("net.2.weight"), 1, wd=self.GW2t)
        K.pack_weights(self.D.p("net.0.weight"), 1, wd=self.DW1t)
        K.pack_weights(self.D.p("net.2.weight"), 1, wd=self.DW2t)

    # ---- pieces
    def _g_fwd(self, z, oh):
        K.copy_cols(z, 0, self.gin, 0, self.zd)
        if self.ld:
            K.copy_cols(oh, 0, self.gin, self.zd, self.ld)
        K.linear_fwd(self.gin, self.G.p("net.0.weight"), self.gh, self.G.p("net.0.bias"), K.ACT_RELU)
        K.linear_fwd(self.gh, self.G.p("net.2.weight"), self.fake, self.G.p("net.2.bias"))

    def _d_fwd(self, rows):
        din, dh, zl = self.din[:rows], self.dh[:rows], self.zl[:rows]
        K.linear_fwd(din, self.D.p("net.0.weight"), dh, self.D.p("net.0.bias"), K.ACT_RELU)
        K.linear_fwd(dh, self.D.p("net.2.weight"), zl, self.D.p("net.2.bias"))

    def _body(self):
        B, ld = self.B, self.ld
        D, G = self.D, self.G
        # ---------------- discriminator step (make_moons_cgan.py:96-111 / make_moons_gan.py:62-74)
        self._g_fwd(self.z1, self.oh1)
        K.copy_cols(self.real, 0, self.din[:B], 0, 2)
        K.copy_cols(self.fake, 0, self.din[B:], 0, 2)
        if ld:
            K.copy_cols(self.real_oh, 0, self.din[:B], 2, ld)
            K.copy_cols(self.oh1, 0, self.din[B:], 2, ld)
        self._d_fwd(2 * B)
        K.gan_loss(self.zl[:B], K.GAN_LOG, 1.0, self.scal[2:3], self.dz[:B], out_aux=self.scal[4:5])
        K.gan_loss(self.zl[B:], K.GAN_LOG, 0.0, self.scal[3:4], self.dz[B:], out_aux=self.scal[5:6])
        K.combine([(1.0, self.scal[2:3]), (1.0, self.scal[3:4])], self.scal[0:1])
        K.linear_wgrad(self.dh, self.dz, self.wsc, D.g("net.2.weight"), D.g("net.2.bias"), self.stat)
        K.linear_dgrad(self.dz, self.DW2t, self.ddh, self.H, act_ref=self.dh, ref_act=K.ACT_RELU)
        K.linear_wgrad(self.din, self.ddh, self.wsc, D.g("net.0.weight"), D.g("net.0.bias"), self.stat)
        D.adam_step(self.lr)
        K.pack_weights(D.p("net.0.weight"), 1, wd=self.DW1t)
        K.pack_weights(D.p("net.2.weight"), 1, wd=self.DW2t)
        # ---------------- generator step (make_moons_cgan.py:114-127 / make_moons_gan.py:77-86)
        self._g_fwd(self.z2, self.oh2)
        K.copy_cols(self.fake, 0, self.din[:B], 0, 2)
        if ld:
            K.copy_cols(self.oh2, 0, self.din[:B], 2, ld)
        self._d_fwd(B)
        K.gan_loss(self.zl[:B], K.GAN_LOG, 1.0, self.scal[1:2], self.dz[:B])
        K.linear_dgrad(self.dz[:B], self.DW2t, self.ddh[:B], self.H, act_ref=self.dh[:B], ref_act=K.ACT_RELU)
        K.linear_dgrad(self.ddh[:B], self.DW1t, self.ddin, self.di)
        K.copy_cols(self.ddin, 0, self.dfake, 0, 2)
        K.linear_wgrad(self.gh, self.dfake, self.wsc, G.g("net.2.weight"), G.g("net.2.bias"), self.stat)
        K.linear_dgrad(self.dfake, self.GW2t, self.dgh, self.H, act_ref=self.gh, ref_act=K.ACT_RELU)
        K.linear_wgrad(self.gin, self.dgh, self.wsc, G.g("net.0.weight"), G.g("net.0.bias"), self.stat)
        G.adam_step(self.lr)
        K.pack_weights(G.p("net.2.weight"), 1, wd=self.GW2t)

    def step(self, real, real_oh, z1, oh1, z2, oh2):
        """Runs one iteration on the injected draws; returns the scalar block."""
        if self.fused:
            # the iteration is ~15 us on the device: keep the host side of the call short (prepared argument tuple,
            # raw data_ptr()s) so the step stays device-bound
            if self._fn is None:
                import ctypes
                from .. import _lib
                self._check = _lib.check
                self._fn = _lib.load().pcg_mlp_gan_step
                vp = ctypes.c_void_p
                self._fn.argtypes = [ctypes.c_int] * 4 + [vp] * 16 + [ctypes.c_float, vp, vp]
                G, D = self.G, self.D
                self._static = tuple(t.data_ptr() for t in (G.data, G.grad, G.m, G.v, G.step, D.data, D.grad, D.m, D.v,
                                                            D.step))
                self._keep = []
            ins = [real, real_oh, z1, oh1, z2, oh2]
            for i, t in enumerate(ins):
                if t is not None and not (t.dtype is torch.float32 and t.is_cuda and t.is_contiguous()):
                    if not t.is_cuda:
                        raise RuntimeError("pcg_b200: inputs must be CUDA tensors (there is no CPU fallback)")
                    ins[i] = t.detach().float().contiguous()
            self._keep = ins            # converted copies stay alive until the next call has been enqueued
            self._check(self._fn(self.B, self.zd, self.ld, self.H, *[None if t is None else t.data_ptr() for t in ins],
                                 *self._static, self.lr, self.scal.data_ptr(), torch.cuda.current_stream().cuda_stream))
            return self.scal
        for dst, src in ((self.real, real), (self.z1, z1), (self.z2, z2), (self.real_oh, real_oh), (self.oh1, oh1),
                         (self.oh2, oh2)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        if not self.use_graph:
            self._body()
            return self.scal
        if self.graph is None:
            # first call runs eagerly once on a snapshot (module loading / attribute setting), then captures
            snap = [t.clone() for t in (self.G.data, self.G.m, self.G.v, self.G.step, self.D.data, self.D.m, self.D.v,
                                        self.D.step)]
            self._body()
            torch.cuda.synchronize()
            for dst, src in zip((self.G.data, self.G.m, self.G.v, self.G.step, self.D.data, self.D.m, self.D.v,
                                 self.D.step), snap):
                dst.copy_(src)
            self.refresh()
            torch.cuda.synchronize()
            self.graph = graphs.capture(self._body)
        self.graph.replay()
        return self.scal


def _plan_for(generator, discriminator, batch, lr, device):
    z_dim_plus = generator.net[0].in_features
    label_dim = discriminator.net[0].in_features - 2
    plan = MlpGanPlan(batch, z_dim_plus - label_dim, label_dim, generator.net[0].out_features, device, lr)
    plan.G.adopt(generator)
    plan.D.adopt(discriminator)
    plan.refresh()
    return plan


def train_gan(X, generator, discriminator, config, device="cuda"):
    """Drop-in for make_moons_gan.py:49-93 (returns (loss_D_values, loss_G_values)); ``n_samples`` must be a
    multiple of ``batch_size`` — the reference draws ``batch_size`` fakes for every real batch (:63) and fails
    otherwise."""
    B = config["batch_size"]
    if len(X) % B:
        raise ValueError("n_samples must be a multiple of batch_size (make_moons_gan.py:63 pairs batch_size fakes "
                         "with every real batch)")
    plan = _plan_for(generator, discriminator, B, config["lr"], device)
    loss_D_values, loss_G_values = [], []
    for epoch in range(config["epochs"]):
        np.random.shuffle(X)
        real_samples = torch.from_numpy(X).float().to(device)
        acc = torch.zeros(8, device=device)
        for real_batch in real_samples.split(B):
            z1 = torch.randn(B, config["z_dim"], device=device)
            z2 = torch.randn(B, config["z_dim"], device=device)
            acc += plan.step(real_batch, None, z1, None, z2, None)
        tot = acc.tolist()
        loss_D_values.append(tot[0])
        loss_G_values.append(tot[1])
    return loss_D_values, loss_G_values


def train_cgan(real_samples, real_labels, generator, discriminator, config, device="cuda"):
    """The training loop of make_moons_cgan.py:83-135 as a function (the reference runs it as the script body).
    Keeps the script's behaviour of drawing the discriminator-step fake labels with ``randint(0, 1)`` (always
    class 0, :98)."""
    B, ld = config["batch_size"], config["label_dim"]
    n = real_samples.shape[0]
    if n % B:
        raise ValueError("n_samples must be a multiple of batch_size (make_moons_cgan.py:97)")
    plan = _plan_for(generator, discriminator, B, config["lr"], device)
    real_samples, real_labels = real_samples.to(device), real_labels.to(device)
    loss_D_values, loss_G_values = [], []
    eye = torch.eye(ld, device=device)
    for epoch in range(config["epochs"]):
        idx = torch.from_numpy(np.random.permutation(n)).to(device)
        real_samples, real_labels = real_samples[idx], real_labels[idx]
        acc = torch.zeros(8, device=device)
        for rb, rl in zip(real_samples.split(B), real_labels.split(B)):
            z1 = torch.randn(B, config["z_dim"], device=device)
            l1 = torch.randint(0, 1, (B,), device=device)
            z2 = torch.randn(B, config["z_dim"], device=device)
            l2 = torch.randint(0, ld, (B,), device=device)
            acc += plan.step(rb, eye[rl], z1, eye[l1], z2, eye[l2])
        tot = acc.tolist()
        loss_D_values.append(tot[0])
        loss_G_values.append(tot[1])
        if epoch % 100 == 0:
            print(f"Epoch [{epoch}/{config['epochs']}], Loss D: {tot[0]:.4f}, Loss G: {tot[1]:.4f}")
    return loss_D_values, loss_G_values
