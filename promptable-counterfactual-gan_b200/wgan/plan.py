"""Native mirror of ``conditional_gan/mnist/mnist_wgan_conditional.py`` (conditional WGAN-GP).

    Hyperparameter                                                                           :20-31
    Generator()  2 embeddings -> 4x ConvTranspose2d (+ BatchNorm2d + ReLU), Tanh             :51-78
    Critic()     3x Conv2d(k3, s2) + InstanceNorm2d(affine) + LeakyReLU(.2), cond. embedding,
                 Linear - LeakyReLU - Linear                                                 :80-108
    critic update with gradient penalty (autograd.grad(create_graph=True))                   :132-154
    generator update every n_critic batches, AdamW(lr 1e-4, betas (0, .9))                   :116-117, :156-168

One iteration is composed from libpcg operators on NHWC fp32 tensors and replayed as a CUDA graph.

The critic update is ONE forward and ONE backward sweep over a batch of 3B samples [real | fake | interpolates]:
InstanceNorm has per-sample statistics, so the three critic calls of the reference are independent rows of one batch, and
``-mean(real) + mean(fake)`` is the upstream gradient (-1/B, +1/B, 0).  The gradient penalty is differentiated by hand:

    forward chain (interpolates)   a_i = conv_i(h_{i-1}),  h_i = lrelu(IN(a_i)),  u = L1 [flat(h_3), emb],  out = L2 lrelu(u)
    gradient chain                 g = d out / d x_hat: gu = L2 * lrelu'(u), gf = gu L1, gh_i -> ga_i = IN_bwd(gh_i) ->
                                   gh_{i-1} = dgrad_i(ga_i) ... g;   gp = lambda * mean_b (||g_b|| - 1)^2
    reverse of the gradient chain  cotangents q: dgrad_i is bilinear, so q_ga_i = conv_i(q_gh_{i-1}) and W_i collects
                                   wgrad(q_gh_{i-1}, ga_i); IN_bwd is differentiated by pcg_instnorm_bwd_bwd, which also
                                   yields the cotangent of a_i (through xhat and 1/sigma); lrelu'' = 0
    those a_i cotangents are INJECTED into the ordinary backward sweep of the forward chain (add_src of IN_bwd on the
    interpolate rows), which carries them to the layers below and to every parameter.

ConvTranspose2d(Cin, Cout) forward is the data gradient of the mirrored convolution Conv2d(Cout -> Cin) and shares its
weight tensor; its input gradient is that convolution's forward (as in pcg_b200.dcgan).
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import graphs
from .. import ops as K


class Hyperparameter:
    """mnist_wgan_conditional.py:20-31 (a plain class: the sizes are plan constants)."""

    def __init__(self, num_classes=10, batchsize=128, num_epochs=20, latent_size=32, n_critic=5, critic_size=1024,
                 generator_size=1024, critic_hidden_size=1024, gp_lambda=10.0, data_path="/mnt/data"):
        self.num_classes, self.batchsize, self.num_epochs, self.latent_size = num_classes, batchsize, num_epochs, latent_size
        self.n_critic, self.critic_size, self.generator_size = n_critic, critic_size, generator_size
        self.critic_hidden_size, self.gp_lambda, self.data_path = critic_hidden_size, gp_lambda, data_path


LR, BETAS, WEIGHT_DECAY = 1e-4, (0.0, 0.9), 1e-2          # :116-117 (AdamW's default decay)
G_K = [4, 3, 4, 4]
G_SP = [(1, 0), (2, 1), (2, 1), (2, 1)]
G_HW = [1, 4, 7, 14, 28]
C_HW = [28, 13, 6, 2]


def g_channels(hp):
    g = hp.generator_size
    return [g, g, g // 2, g // 4, 1]


def c_channels(hp):
    c = hp.critic_size
    return [1, c // 4, c // 2, c]


def g_shapes(hp):
    g, ch = hp.generator_size, g_channels(hp)
    s = OrderedDict()
    s["latent_embedding.0.weight"], s["latent_embedding.0.bias"] = (g // 2, hp.latent_size), (g // 2,)
    s["condition_embedding.0.weight"], s["condition_embedding.0.bias"] = (g // 2, hp.num_classes), (g // 2,)
    for i in range(4):
        s[f"tcnn.{3 * i}.weight"], s[f"tcnn.{3 * i}.bias"] = (ch[i], ch[i + 1], G_K[i], G_K[i]), (ch[i + 1],)
        if i < 3:
            s[f"tcnn.{3 * i + 1}.weight"], s[f"tcnn.{3 * i + 1}.bias"] = (ch[i + 1],), (ch[i + 1],)
    return s


def c_shapes(hp):
    c, ch = hp.critic_size, c_channels(hp)
    s = OrderedDict()
    s["condition_embedding.0.weight"], s["condition_embedding.0.bias"] = (c * 4, hp.num_classes), (c * 4,)
    for i in range(3):
        s[f"cnn_net.{3 * i}.weight"], s[f"cnn_net.{3 * i}.bias"] = (ch[i + 1], ch[i], 3, 3), (ch[i + 1],)
        s[f"cnn_net.{3 * i + 1}.weight"], s[f"cnn_net.{3 * i + 1}.bias"] = (ch[i + 1],), (ch[i + 1],)
    s["Critic_net.0.weight"], s["Critic_net.0.bias"] = (hp.critic_hidden_size, c * 8), (hp.critic_hidden_size,)
    s["Critic_net.2.weight"], s["Critic_net.2.bias"] = (1, hp.critic_hidden_size), (1,)
    return s


class Generator(nn.Module):
    def __init__(self, hp=None):
        super().__init__()
        hp = self.hp = hp or Hyperparameter()
        g, ch = hp.generator_size, g_channels(hp)
        self.latent_embedding = nn.Sequential(nn.Linear(hp.latent_size, g // 2))
        self.condition_embedding = nn.Sequential(nn.Linear(hp.num_classes, g // 2))
        layers = []
        for i in range(4):
            layers.append(nn.ConvTranspose2d(ch[i], ch[i + 1], G_K[i], *G_SP[i]))
            layers += [nn.BatchNorm2d(ch[i + 1]), nn.ReLU(inplace=True)] if i < 3 else [nn.Tanh()]
        self.tcnn = nn.Sequential(*layers)

    def forward(self, latent, condition):
        """[B, latent] , one-hot [B, classes] -> [B,1,28,28]; train mode uses (and moves) the batch statistics."""
        return _forward_plan(self, latent.shape[0], "G").g_forward(latent, condition, self.training)


class Critic(nn.Module):
    def __init__(self, hp=None):
        super().__init__()
        hp = self.hp = hp or Hyperparameter()
        c, ch = hp.critic_size, c_channels(hp)
        self.condition_embedding = nn.Sequential(nn.Linear(hp.num_classes, c * 4))
        layers = []
        for i in range(3):
            layers += [nn.Conv2d(ch[i], ch[i + 1], 3, 2), nn.InstanceNorm2d(ch[i + 1], affine=True),
                       nn.LeakyReLU(0.2, inplace=True)]
        self.cnn_net = nn.Sequential(*layers, nn.Flatten())
        self.Critic_net = nn.Sequential(nn.Linear(c * 8, hp.critic_hidden_size), nn.LeakyReLU(0.2, inplace=True),
                                        nn.Linear(hp.critic_hidden_size, 1))

    def forward(self, image, condition):
        return _forward_plan(self, image.shape[0], "C").c_forward(image, condition)


def _forward_plan(module, batch, which):
    cache = module.__dict__.setdefault("_pcg_plans", {})
    p = cache.get(batch)
    if p is not None and not (p.G if which == "G" else p.C).aliases(module):
        p = None
    if p is None:
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("pcg_b200: modules must live on a CUDA device (there is no CPU fallback)")
        p = WganGpPlan(module.hp, batch, dev, use_graph=False)
        (p.adopt_g if which == "G" else p.adopt_c)(module)
        cache.clear()
        cache[batch] = p
    p.refresh()
    return p


class WganGpPlan:
    def __init__(self, hp, batch, device, use_graph=True, tensor_cores=None, operand_terms=None, share=None):
        """tensor_cores: the 64-multiple-channel convolutions / linear layers on the tcgen05 kernels with bf16x3 operands
        (None = follow PCG_PRECISION, default bf16 -> on; False = exact fp32 on the CUDA cores).  share: another plan
        (other batch size) whose parameters, optimizer state and BatchNorm buffers this one uses."""
        self.hp, self.B = hp, batch
        self.tc = (os.environ.get("PCG_PRECISION", "bf16") != "fp32") if tensor_cores is None else bool(tensor_cores)
        # operand precision of the tensor-core products: 3 = bf16x3 (fp32-equivalent products; default here - the penalty is
        # a second derivative, with plain bf16 operands (1) the first critic layers' gradients are off by 10-30 %,
        # tests/test_wgan_gpu.py), 1 = plain bf16 (1.4x faster, PCG_TC_TERMS=1)
        self.terms = int(os.environ.get("PCG_TC_TERMS", "3")) if operand_terms is None else int(operand_terms)
        dev = self.dev = torch.device(device)
        B, B3 = batch, 3 * batch
        c, g, Hd, nc = hp.critic_size, hp.generator_size, hp.critic_hidden_size, hp.num_classes
        gch, cch = self.gch, self.cch = g_channels(hp), c_channels(hp)
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        if share is not None:
            self.G, self.C, self.C_grad2, self.g_bn, self.lr_dev = share.G, share.C, share.C_grad2, share.g_bn, share.lr_dev
            self.g_wf, self.g_wd, self.c_wf, self.c_wd, self.l1t = share.g_wf, share.g_wd, share.c_wf, share.c_wd, share.l1t
            self.g_wc, self.c_wc = share.g_wc, share.c_wc
        else:
            self.G, self.C = K.FlatParams(list(g_shapes(hp).items()), dev), K.FlatParams(list(c_shapes(hp).items()), dev)
            self.C_grad2 = torch.zeros_like(self.C.grad)       # second-order (penalty) terms of the gradient chain
            self.g_bn = [dict(rm=z(gch[i + 1]), rv=torch.ones(gch[i + 1], device=dev),
                              nbt=torch.zeros((), dtype=torch.int64, device=dev), st=K.BNState(gch[i + 1], dev))
                         for i in range(3)]
            self.lr_dev = torch.full((1,), LR, device=dev)
            self.g_wf = [z(gch[i] * gch[i + 1] * G_K[i] ** 2) for i in range(4)]
            self.g_wd = [z(gch[i] * gch[i + 1] * G_K[i] ** 2) for i in range(4)]
            self.c_wf = [z(cch[i] * cch[i + 1] * 9) for i in range(3)]
            self.c_wd = [z(cch[i] * cch[i + 1] * 9) for i in range(3)]
            self.l1t = z(8 * c * Hd)                           # Critic_net.0.weight transposed [8c][hidden]
            # forward weights of the four parity-class convolutions of a stride-2 data gradient (tensor-core mode, _dgrad)
            self.g_wc = [z(4, gch[i] * 4 * gch[i + 1]) if i == 1 else None for i in range(4)]
            self.c_wc = [z(4, cch[i] * 4 * cch[i + 1]) if i > 0 else None for i in range(3)]
        # mirrored-convolution geometry of the generator's transposed convolutions
        self.g_geom = [(B, G_HW[i + 1], G_HW[i + 1], gch[i + 1], gch[i], G_K[i], *G_SP[i]) for i in range(4)]
        # ---- inputs
        self.real, self.noise, self.noise_g = z(B, 784), z(B, hp.latent_size), z(B, hp.latent_size)
        self.alpha = z(B, 784)                                   # alpha[b] broadcast over the image
        self.labels = torch.zeros(B, dtype=torch.int64, device=dev)
        self.labels_g = torch.zeros(B, dtype=torch.int64, device=dev)
        self.cond3, self.cond_g = z(B3, nc), z(B, nc)
        # ---- generator activations
        self.el, self.ec, self.e = z(B, g // 2), z(B, g // 2), z(B, g)
        self.gy = [z(B, G_HW[i + 1], G_HW[i + 1], gch[i + 1]) for i in range(4)]
        self.ga = [z(B, G_HW[i + 1], G_HW[i + 1], gch[i + 1]) for i in range(4)]
        self.gd = [z(B, G_HW[i + 1], G_HW[i + 1], gch[i + 1]) for i in range(4)]
        self.gdy = [z(B, G_HW[i + 1], G_HW[i + 1], gch[i + 1]) for i in range(4)]
        self.de, self.del_, self.dec = z(B, g), z(B, g // 2), z(B, g // 2)
        # ---- critic: one 3B batch [real | fake | interpolates]
        act = lambda n: [z(n, C_HW[i + 1], C_HW[i + 1], cch[i + 1]) for i in range(3)]  # noqa: E731
        stat = lambda n: [z(n, cch[i + 1]) for i in range(3)]  # noqa: E731
        self.x3, self.tmp = z(B3, 784), z(B, 784)
        self.a, self.h, self.mean, self.rstd = act(B3), act(B3), stat(B3), stat(B3)
        self.vc, self.f, self.v, self.out = z(B3, 4 * c), z(B3, 8 * c), z(B3, Hd), z(B3, 1)
        self.dout3 = torch.cat([torch.full((B, 1), -1.0 / B), torch.full((B, 1), 1.0 / B), torch.zeros(B, 1)]).to(dev)
        self.dout_g = torch.full((B, 1), -1.0 / B, device=dev)
        self.du, self.df, self.dvc = z(B3, Hd), z(B3, 8 * c), z(B3, 4 * c)
        self.dh, self.da, self.gpart, self.bpart = act(B3), act(B3), stat(B3), stat(B3)
        self.inj = act(B3)                  # a_i cotangents of the penalty (interpolate rows; the other rows stay zero)
        self.dx = z(B3, 784)
        # ---- gradient chain of the interpolates and its reverse
        self.ones = torch.ones(B, 1, device=dev)
        self.gu, self.gf, self.gh, self.ga_c = z(B, Hd), z(B, 8 * c), act(B), act(B)
        self.gimg, self.gbar, self.norms = z(B, 784), z(B, 784), z(B)
        self.q_ga, self.q_gh, self.g2part = act(B), act(B), stat(B)
        self.qf, self.q_gu, self.q_gv = z(B, 8 * c), z(B, Hd), z(B, Hd)
        self.scal = z(8)                    # 0 critic_loss, 1 mean D(real), 2 mean D(fake), 3 penalty, 4 generator_loss
        self._scratch = {}
        self.use_graph, self.graphs = use_graph, {}
        self.refresh()

    # ------------------------------------------------------------------ scratch per call site
    def _ws(self, key, *geom):
        t = self._scratch.get(("w", key))
        if t is None:
            t = self._scratch[("w", key)] = K.conv_wgrad_scratch(*geom, self.dev)
        return t

    def _ss(self, key, C):
        t = self._scratch.get(("s", key))
        if t is None:
            t = self._scratch[("s", key)] = K.stat_scratch(C, self.dev)
        return t

    def _buf(self, key, *shape):
        t = self._scratch.get(("b", key))
        if t is None:
            t = self._scratch[("b", key)] = torch.zeros(*shape, device=self.dev)
        return t

    def _lin_wgrad(self, key, x, dy, dw, db):
        """dw[N][K] = dy^T x.  Tensor-core mode: as the forward product of the two TRANSPOSED activations (rows = N,
        reduction over the batch), which is what the tcgen05 forward kernel computes."""
        Bn, Kd = x.shape
        N = dy.shape[1]
        if self.tc and Bn % 64 == 0 and Kd % 64 == 0 and N >= 64:
            xT, dyT = self._buf(("xT", key), 1, Kd * Bn), self._buf(("dyT", key), 1, N * Bn)
            K.flatten_nchw(x, 1, Bn, Kd, xT, Kd * Bn, 0)              # [Bn][Kd] -> [Kd][Bn]
            K.flatten_nchw(dy, 1, Bn, N, dyT, N * Bn, 0)
            K.conv_fprop(dyT, N, 1, 1, Bn, xT, Kd, 1, 1, 0, dw)
        else:
            K.conv_wgrad(x, dy, Bn, 1, 1, Kd, N, 1, 1, 0, self._ws(key, Bn, 1, 1, Kd, N, 1, 1, 0), dw)
        if db is not None:
            K.colsum(dy, self._ss(key, N), db)

    def _lin_dgrad(self, dy, w_t, dx, Kd):
        """dx[B][K] = dy[B][N] W[N][K] with w_t = W^T [K][N]: in tensor-core mode the forward product with W^T as weight."""
        if self.tc:
            K.linear_fwd(dy, w_t, dx)
        else:
            K.linear_dgrad(dy, w_t, dx, Kd)

    def _dgrad(self, key, dy, geom, wd, wc, out):
        """Data gradient of the convolution ``geom`` = (N, H, W, Cin, Cout, k, stride, pad).  Tensor-core mode, for the
        geometries the native tcgen05 data-gradient kernel (k4, s2, p1, even sizes) does not cover, as FORWARD
        convolutions, which the tcgen05 forward kernel takes in any geometry:
          stride 2 (k3; any padding, odd sizes): one stride-1 2x2 convolution of the gradient per parity class of the input
            position - all four as ONE Cout -> 4 Cin layer (weights ``wc``, pcg_pack_dgrad_classes) - interleaved by
            pcg_parity_interleave: 16 tap products per output pixel pair against 36 for the zero-dilated form (pcg_dilate);
          a full-window layer (ConvTranspose2d on the 1x1 latent): a 1x1 product with wd's rows, (ci, tap) -> (tap, ci)."""
        N, H, W, Cin, Cout, k, stride, pad = geom
        native = k == 4 and stride == 2 and pad == 1 and H % 2 == 0
        Ho = (H + 2 * pad - k) // stride + 1
        ok = self.tc and Cin % 64 == 0 and Cout % 64 == 0 and not native
        if ok and stride == 2 and wc is not None:
            cls = self._buf(("cls", key), N, Ho + 1, Ho + 1, 4 * Cin)        # the four classes as ONE Cout -> 4 Cin layer
            K.conv_fprop(dy, N, Ho, Ho, Cout, wc, 4 * Cin, 2, 1, 1, cls)
            K.parity_interleave(cls, N, Ho + 1, Ho + 1, Cin, pad, H, W, out, stacked=True)
        elif ok and stride == 1 and pad == 0 and Ho == 1:
            tmp = self._buf(("fw", key), N, Cin * k * k)
            K.conv_fprop(dy, N, 1, 1, Cout, wd, Cin * k * k, 1, 1, 0, tmp)
            K.flatten_nchw(tmp, N, Cin, k * k, out.view(N, k * k * Cin), k * k * Cin, 0)
        else:
            K.conv_dgrad(dy, N, H, W, Cin, wd, Cout, k, stride, pad, out)

    # ------------------------------------------------------------------ binding
    def _adopt_bn(self, module):
        bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm2d)]
        for b, m in zip(self.g_bn, bns):
            for key, name in (("rm", "running_mean"), ("rv", "running_var"), ("nbt", "num_batches_tracked")):
                b[key].copy_(getattr(m, name))
                m._buffers[name] = b[key]

    def adopt_g(self, module):
        self.G.adopt(module)
        self._adopt_bn(module)
        self.refresh()

    def adopt_c(self, module):
        self.C.adopt(module)
        self.refresh()

    def refresh(self):
        self._pack_g()
        self._pack_c()

    def _pack_g(self):
        for i in range(4):
            # ConvT weight [Cin_T][Cout_T][k][k] == mirrored conv weight OIHW with O = Cin_T, I = Cout_T
            K.pack_weights(self.G.p(f"tcnn.{3 * i}.weight"), G_K[i], wf=self.g_wf[i], wd=self.g_wd[i])
            if self.tc and self.g_wc[i] is not None:
                K.pack_dgrad_classes(self.G.p(f"tcnn.{3 * i}.weight"), G_K[i], self.g_wc[i])

    def _pack_c(self):
        for i in range(3):
            K.pack_weights(self.C.p(f"cnn_net.{3 * i}.weight"), 3, wf=self.c_wf[i], wd=self.c_wd[i])
            if self.tc and self.c_wc[i] is not None:
                K.pack_dgrad_classes(self.C.p(f"cnn_net.{3 * i}.weight"), 3, self.c_wc[i])
        K.pack_weights(self.C.p("Critic_net.0.weight"), 1, wd=self.l1t)

    # ------------------------------------------------------------------ generator
    def _g_fwd(self, noise, cond, training=True):
        G, g = self.G, self.hp.generator_size
        K.linear_fwd(noise, G.p("latent_embedding.0.weight"), self.el, bias=G.p("latent_embedding.0.bias"))
        K.linear_fwd(cond, G.p("condition_embedding.0.weight"), self.ec, bias=G.p("condition_embedding.0.bias"))
        K.copy_cols(self.el, 0, self.e, 0, g // 2)                                   # torch.cat(dim=1), :77
        K.copy_cols(self.ec, 0, self.e, g // 2, g // 2)
        x = self.e
        for i, geom in enumerate(self.g_geom):
            C = self.gch[i + 1]
            self._dgrad(("g", i), x, geom, self.g_wd[i], self.g_wc[i], self.gy[i])  # ConvTranspose2d forward
            if i < 3:
                K.bias_act(self.gy[i], C, G.p(f"tcnn.{3 * i}.bias"), self.gy[i])
                bn, nm = self.g_bn[i], f"tcnn.{3 * i + 1}"
                if training:
                    K.bn_train_fwd(self.gy[i], self.B * G_HW[i + 1] ** 2, C, G.p(nm + ".weight"), G.p(nm + ".bias"), bn["rm"],
                                   bn["rv"], bn["nbt"], bn["st"], self.ga[i], act=K.ACT_RELU)
                else:
                    K.bn_eval(self.gy[i], G.p(nm + ".weight"), G.p(nm + ".bias"), bn["rm"], bn["rv"], self.ga[i])
                    K.unary(self.ga[i], K.RELU, self.ga[i])
            else:
                K.bias_act(self.gy[i], C, G.p(f"tcnn.{3 * i}.bias"), self.ga[i], tanh_out=True)
            x = self.ga[i]

    def _g_bwd(self, dfake, noise, cond):
        """dfake (gradient wrt the tanh output) -> every generator parameter gradient."""
        G, g = self.G, self.hp.generator_size
        K.unary_bwd(dfake, self.ga[3], K.TANH, self.gdy[3])
        for i in range(3, -1, -1):
            geom = self.g_geom[i]
            xin = self.e if i == 0 else self.ga[i - 1]
            K.conv_wgrad(self.gdy[i], xin, *geom, self._ws(("g", i), *geom), G.g(f"tcnn.{3 * i}.weight"))
            K.colsum(self.gdy[i], self._ss(("g", i), self.gch[i + 1]), G.g(f"tcnn.{3 * i}.bias"))
            dst = self.de if i == 0 else self.gd[i - 1]
            K.conv_fprop(self.gdy[i], *geom[:4], self.g_wf[i], *geom[4:], dst)       # ConvT input gradient
            if i == 0:
                break
            nm, bn = f"tcnn.{3 * (i - 1) + 1}", self.g_bn[i - 1]
            K.bn_train_bwd(self.gd[i - 1], self.gy[i - 1], self.B * G_HW[i] ** 2, self.gch[i], G.p(nm + ".weight"), bn["st"],
                           self.gdy[i - 1], G.g(nm + ".weight"), G.g(nm + ".bias"), act=K.ACT_RELU)
        K.copy_cols(self.de, 0, self.del_, 0, g // 2)
        K.copy_cols(self.de, g // 2, self.dec, 0, g // 2)
        self._lin_wgrad("gl", noise, self.del_, G.g("latent_embedding.0.weight"), G.g("latent_embedding.0.bias"))
        self._lin_wgrad("gc", cond, self.dec, G.g("condition_embedding.0.weight"), G.g("condition_embedding.0.bias"))

    # ------------------------------------------------------------------ critic
    def _c_fwd(self, N, cond):
        """Rows [0, N) of x3 through the critic (N = 3B in the critic update, B elsewhere); scores land in out[:N]."""
        Cp, c = self.C, self.hp.critic_size
        x = self.x3[:N]
        for i in range(3):
            Hh, Ci, Co = C_HW[i], self.cch[i], self.cch[i + 1]
            K.conv_fprop(x, N, Hh, Hh, Ci, self.c_wf[i], Co, 3, 2, 0, self.a[i][:N], bias=Cp.p(f"cnn_net.{3 * i}.bias"))
            K.instnorm_fwd(self.a[i][:N], N, C_HW[i + 1] ** 2, Co, Cp.p(f"cnn_net.{3 * i + 1}.weight"),
                           Cp.p(f"cnn_net.{3 * i + 1}.bias"), self.h[i][:N], self.mean[i][:N], self.rstd[i][:N],
                           act=K.ACT_LRELU, slope=0.2)
            x = self.h[i][:N]
        K.linear_fwd(cond[:N], Cp.p("condition_embedding.0.weight"), self.vc[:N], bias=Cp.p("condition_embedding.0.bias"))
        K.flatten_nchw(self.h[2][:N], N, 4, c, self.f[:N], 8 * c, 0)                  # nn.Flatten of NCHW + torch.cat :104-105
        K.copy_cols(self.vc[:N], 0, self.f[:N], 4 * c, 4 * c)
        K.linear_fwd(self.f[:N], Cp.p("Critic_net.0.weight"), self.v[:N], bias=Cp.p("Critic_net.0.bias"), act=K.ACT_LRELU,
                     slope=0.2)
        K.linear_fwd(self.v[:N], Cp.p("Critic_net.2.weight"), self.out[:N], bias=Cp.p("Critic_net.2.bias"))

    def _c_bwd(self, N, dout, cond, want_wgrad, want_dx, inject):
        """Backward of rows [0, N) from the score gradient ``dout``; parameter gradients into the main arena."""
        Cp, c, Hd = self.C, self.hp.critic_size, self.hp.critic_hidden_size
        g = Cp.g
        l2t = Cp.p("Critic_net.2.weight").view(Hd, 1)
        K.linear_dgrad(dout[:N], l2t, self.du[:N], Hd, act_ref=self.v[:N], ref_act=K.ACT_LRELU, ref_slope=0.2)
        if want_wgrad:
            self._lin_wgrad(("l2", N), self.v[:N], dout[:N], g("Critic_net.2.weight"), g("Critic_net.2.bias"))
            self._lin_wgrad(("l1", N), self.f[:N], self.du[:N], g("Critic_net.0.weight"), g("Critic_net.0.bias"))
        self._lin_dgrad(self.du[:N], self.l1t.view(8 * c, Hd), self.df[:N], 8 * c)
        if want_wgrad:
            K.copy_cols(self.df[:N], 4 * c, self.dvc[:N], 0, 4 * c)
            self._lin_wgrad(("ce", N), cond[:N], self.dvc[:N], g("condition_embedding.0.weight"),
                            g("condition_embedding.0.bias"))
        K.flatten_nchw(self.df[:N], N, 4, c, self.dh[2][:N], 8 * c, 0, inverse=True)
        for i in range(2, -1, -1):
            Hh, Ci, Co, P = C_HW[i], self.cch[i], self.cch[i + 1], C_HW[i + 1] ** 2
            K.instnorm_bwd(self.dh[i][:N], self.a[i][:N], self.mean[i][:N], self.rstd[i][:N],
                           Cp.p(f"cnn_net.{3 * i + 1}.weight"), N, P, Co, self.da[i][:N], act_ref=self.h[i][:N],
                           act=K.ACT_LRELU, slope=0.2, add_src=inject[i][:N] if inject is not None else None,
                           dgamma_part=self.gpart[i][:N] if want_wgrad else None,
                           dbeta_part=self.bpart[i][:N] if want_wgrad else None)
            xin = self.x3[:N] if i == 0 else self.h[i - 1][:N]
            if want_wgrad:
                K.colsum(self.gpart[i][:N], self._ss(("ing", i), Co), g(f"cnn_net.{3 * i + 1}.weight"))
                K.colsum(self.bpart[i][:N], self._ss(("inb", i), Co), g(f"cnn_net.{3 * i + 1}.bias"))
                geom = (N, Hh, Hh, Ci, Co, 3, 2, 0)
                K.conv_wgrad(xin, self.da[i][:N], *geom, self._ws(("c", i, N), *geom), g(f"cnn_net.{3 * i}.weight"))
                K.colsum(self.da[i][:N], self._ss(("cb", i), Co), g(f"cnn_net.{3 * i}.bias"))
            if i > 0:
                self._dgrad(("c", i, N), self.da[i][:N], (N, Hh, Hh, Ci, Co, 3, 2, 0), self.c_wd[i], self.c_wc[i],
                            self.dh[i - 1][:N])
            elif want_dx:
                K.conv_dgrad(self.da[0][:N], N, Hh, Hh, Ci, self.c_wd[0], Co, 3, 2, 0, self.dx[:N])

    def _penalty(self):
        """Gradient chain of the interpolate rows, the penalty, and the reverse of the chain (:146-150)."""
        B, Cp, c, Hd = self.B, self.C, self.hp.critic_size, self.hp.critic_hidden_size
        s = slice(2 * B, 3 * B)
        g2 = lambda n: Cp._view(self.C_grad2, n)  # noqa: E731
        gam = lambda i: Cp.p(f"cnn_net.{3 * i + 1}.weight")  # noqa: E731
        l2t = Cp.p("Critic_net.2.weight").view(Hd, 1)
        # d out / d x_hat, grad_outputs = ones (:146)
        K.linear_dgrad(self.ones, l2t, self.gu, Hd, act_ref=self.v[s], ref_act=K.ACT_LRELU, ref_slope=0.2)
        self._lin_dgrad(self.gu, self.l1t.view(8 * c, Hd), self.gf, 8 * c)
        K.flatten_nchw(self.gf, B, 4, c, self.gh[2], 8 * c, 0, inverse=True)
        for i in range(2, -1, -1):
            Hh, Ci, Co, P = C_HW[i], self.cch[i], self.cch[i + 1], C_HW[i + 1] ** 2
            K.instnorm_bwd(self.gh[i], self.a[i][s], self.mean[i][s], self.rstd[i][s], gam(i), B, P, Co, self.ga_c[i],
                           act_ref=self.h[i][s], act=K.ACT_LRELU, slope=0.2)
            self._dgrad(("p", i), self.ga_c[i], (B, Hh, Hh, Ci, Co, 3, 2, 0), self.c_wd[i], self.c_wc[i],
                        self.gh[i - 1] if i > 0 else self.gimg)
        K.gp_penalty(self.gimg, B, 784, self.hp.gp_lambda, self.scal[3:4], self.gbar, self.norms)
        # reverse: cotangents of the chain's intermediates, second-order parameter terms into C_grad2
        q = self.gbar
        for i in range(3):
            Hh, Ci, Co, P = C_HW[i], self.cch[i], self.cch[i + 1], C_HW[i + 1] ** 2
            geom = (B, Hh, Hh, Ci, Co, 3, 2, 0)
            K.conv_fprop(q, B, Hh, Hh, Ci, self.c_wf[i], Co, 3, 2, 0, self.q_ga[i])
            K.conv_wgrad(q, self.ga_c[i], *geom, self._ws(("q", i), *geom), g2(f"cnn_net.{3 * i}.weight"))
            K.instnorm_bwd_bwd(self.q_ga[i], self.gh[i], self.a[i][s], self.mean[i][s], self.rstd[i][s], gam(i), B, P, Co,
                               self.q_gh[i], self.inj[i][s], act_ref=self.h[i][s], act=K.ACT_LRELU, slope=0.2,
                               dgamma_part=self.g2part[i])
            K.colsum(self.g2part[i], self._ss(("q", i), Co), g2(f"cnn_net.{3 * i + 1}.weight"))
            q = self.q_gh[i]
        K.flatten_nchw(self.q_gh[2], B, 4, c, self.qf, 8 * c, 0)          # the embedding half of qf stays zero
        K.linear_fwd(self.qf, Cp.p("Critic_net.0.weight"), self.q_gu)
        self._lin_wgrad("ql1", self.qf, self.gu, g2("Critic_net.0.weight"), None)
        K.unary_bwd(self.q_gu, self.v[s], K.LRELU, self.q_gv, 0.2)
        K.colsum(self.q_gv, self._ss("ql2", Hd), g2("Critic_net.2.weight").view(-1))

    # ------------------------------------------------------------------ one iteration
    def _tc(self, fn):
        K.set_conv_tensor_cores(self.tc)
        prev = K.set_conv_tensor_core_terms(self.terms)
        # bf16 conversions of unchanged operands are reused within this body (ops.set_operand_cache); nothing survives
        # from before it: the step inputs were written by torch
        prev_cache = K.set_operand_cache(self.tc and os.environ.get("PCG_OPERAND_CACHE", "1") != "0")
        K.operand_cache_clear()
        try:
            fn()
        finally:
            K.set_conv_tensor_cores(False)
            K.set_conv_tensor_core_terms(prev)
            K.set_operand_cache(prev_cache)

    def _critic_grads(self):
        B, nc = self.B, self.hp.num_classes
        K.onehot(self.labels, nc, self.cond3[:B])
        K.copy_cols(self.cond3[:B], 0, self.cond3[B:2 * B], 0, nc)
        K.copy_cols(self.cond3[:B], 0, self.cond3[2 * B:], 0, nc)
        self._g_fwd(self.noise, self.cond3[:B])                              # under no_grad in the reference (:140)
        fake = self.ga[3].view(B, 784)
        K.unary(self.real, K.COPY, self.x3[:B])
        K.unary(fake, K.COPY, self.x3[B:2 * B])
        K.binary(self.real, fake, K.ADD, self.tmp, 1.0, -1.0)                # alpha * real + (1 - alpha) * fake (:145)
        K.film_fwd(self.alpha, self.tmp, fake, self.x3[2 * B:])
        self._c_fwd(3 * B, self.cond3)
        K.reduce_scalar(self.out[:B], self.scal[1:2], scale=1.0 / B)
        K.reduce_scalar(self.out[B:2 * B], self.scal[2:3], scale=1.0 / B)
        self._penalty()
        self._c_bwd(3 * B, self.dout3, self.cond3, True, False, self.inj)
        K.binary(self.C.grad, self.C_grad2, K.ADD, self.C.grad)
        K.combine([(-1.0, self.scal[1:2]), (1.0, self.scal[2:3]), (1.0, self.scal[3:4])], self.scal[0:1])

    def _critic_update(self):
        Cp = self.C
        K.adamw(Cp.data, Cp.grad, Cp.m, Cp.v, Cp.step, self.lr_dev, BETAS[0], BETAS[1], 1e-8, WEIGHT_DECAY)
        self._pack_c()

    def _generator_grads(self):
        B, nc = self.B, self.hp.num_classes
        K.onehot(self.labels_g, nc, self.cond_g)
        self._g_fwd(self.noise_g, self.cond_g)
        K.unary(self.ga[3].view(B, 784), K.COPY, self.x3[:B])
        self._c_fwd(B, self.cond_g)
        K.reduce_scalar(self.out[:B], self.scal[4:5], scale=-1.0 / B)         # generator_loss = -mean D(fake) (:164)
        self._c_bwd(B, self.dout_g, self.cond_g, False, True, None)
        self._g_bwd(self.dx[:B], self.noise_g, self.cond_g)

    def _generator_update(self):
        G = self.G
        K.adamw(G.data, G.grad, G.m, G.v, G.step, self.lr_dev, BETAS[0], BETAS[1], 1e-8, WEIGHT_DECAY)
        self._pack_g()

    def _body(self, with_generator):
        def run():
            self._critic_grads()
            self._critic_update()
            if with_generator:
                self._generator_grads()
                self._generator_update()
        self._tc(run)

    def _state(self):
        t = [self.G.data, self.G.m, self.G.v, self.G.step, self.C.data, self.C.m, self.C.v, self.C.step, self.scal]
        for b in self.g_bn:
            t += [b["rm"], b["rv"], b["nbt"]]
        return t

    def load_inputs(self, real, labels, noise, alpha, labels_g=None, noise_g=None):
        B = self.B
        self.real.copy_(real.reshape(B, 784), non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.noise.copy_(noise.reshape(B, -1), non_blocking=True)
        self.alpha.copy_(alpha.reshape(B, 1).expand(B, 784), non_blocking=True)
        if labels_g is not None:
            self.labels_g.copy_(labels_g, non_blocking=True)
            self.noise_g.copy_(noise_g.reshape(B, -1), non_blocking=True)

    def step(self, real, labels, noise, alpha, labels_g=None, noise_g=None):
        """One batch of the loop body: the critic update, and - when ``labels_g`` / ``noise_g`` are given (the reference
        does it for batch_idx % n_critic == 0) - the generator update.  real [B,1,28,28], labels int64 [B], noise
        [B, latent], alpha [B,1].  Returns the scalar block (see ``scal``)."""
        with_g = labels_g is not None
        self.load_inputs(real, labels, noise, alpha, labels_g, noise_g)
        self.run(with_g)
        return self.scal

    def run(self, with_generator):
        """The iteration on the inputs already staged in the plan's static buffers."""
        if not self.use_graph:
            self._body(with_generator)
            return
        gr = self.graphs.get(with_generator)
        if gr is None:
            snap = [t.clone() for t in self._state()]
            self._body(with_generator)                  # eager dry run: sizes library scratch outside the capture
            torch.cuda.synchronize()
            for dst, src in zip(self._state(), snap):
                dst.copy_(src)
            self.refresh()
            torch.cuda.synchronize()
            gr = self.graphs[with_generator] = graphs.capture(lambda: self._body(with_generator))
        gr.replay()

    # ------------------------------------------------------------------ module forwards
    def g_forward(self, latent, condition, training):
        with torch.no_grad():
            B = self.B
            self.noise.copy_(latent.reshape(B, -1))
            self.cond_g.copy_(condition.reshape(B, -1).float())
            self._tc(lambda: self._g_fwd(self.noise, self.cond_g, training))
            return self.ga[3].view(B, 1, 28, 28).clone()

    def c_forward(self, image, condition):
        with torch.no_grad():
            B = self.B
            self.x3[:B].copy_(image.reshape(B, 784))
            self.cond_g.copy_(condition.reshape(B, -1).float())
            self._tc(lambda: self._c_fwd(B, self.cond_g))
            return self.out[:B].clone()


def train_wgan_gp(generator, critic, dataloader, hp, device="cuda", log=print):
    """The training loop of mnist_wgan_conditional.py:128-190 as a function.  Returns (generator_losses, critic_losses),
    the per-epoch averages the script plots.  The draws of :139 / :144 / :160-161 come from torch's CUDA generator."""
    plans = {}
    generator_losses, critic_losses = [], []
    for epoch in range(hp.num_epochs):
        acc = torch.zeros(8, device=device)
        n = 0
        for batch_idx, data in enumerate(dataloader):
            real = data[0].to(device, non_blocking=True).float()
            labels = data[1].to(device, non_blocking=True).long()
            b = real.size(0)
            plan = plans.get(b)
            if plan is None:
                first = next(iter(plans.values()), None)
                plan = plans[b] = WganGpPlan(hp, b, device, share=first)
                if first is None:
                    plan.adopt_g(generator)
                    plan.adopt_c(critic)
            noise = torch.randn((b, hp.latent_size), device=device)
            alpha = torch.rand((b, 1), device=device)
            if batch_idx % hp.n_critic == 0:
                labels_g = torch.randint(hp.num_classes, size=[b], device=device)
                noise_g = torch.randn((b, hp.latent_size), device=device)
                sc = plan.step(real, labels, noise, alpha, labels_g, noise_g)
            else:
                sc = plan.step(real, labels, noise, alpha)
            acc += sc            # scal[4] keeps the last generator loss, as epoch_g_losses does at :178
            n += 1
        tot = acc.tolist()
        generator_losses.append(tot[4] / max(n, 1))
        critic_losses.append(tot[0] / max(n, 1))
        log(f"[{epoch + 1:>2}/{hp.num_epochs}]  Avg D Loss: {critic_losses[-1]:.4f}  Avg G Loss: {generator_losses[-1]:.4f}")
    return generator_losses, critic_losses
