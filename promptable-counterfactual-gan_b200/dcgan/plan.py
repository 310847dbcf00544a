"""Native mirror of ``dconv_gan/mnist/mnist_dcgan.py``.

    Generator()      5x ConvTranspose2d(k4) + BatchNorm2d + ReLU, Tanh      mnist_dcgan.py:72-93
    Discriminator()  5x Conv2d(k4) + BatchNorm2d + LeakyReLU(.2), Sigmoid   mnist_dcgan.py:96-116
    weights_init                                                             mnist_dcgan.py:63-69
    loop body (D on real, D on fake.detach(), Adam(.5,.999) on D, D(fake) -> G, Adam on G)   :147-175

One iteration is composed from libpcg operators on NHWC fp32 tensors and replayed as a CUDA graph.
ConvTranspose2d(Cin, Cout) forward is the data gradient of the mirrored convolution Conv2d(Cout -> Cin) and
shares its weight tensor ([Cin][Cout][k][k] in both views); its input gradient is that convolution's forward.
BatchNorm runs in train mode in all three discriminator passes (separate batch statistics for real and fake,
running buffers updated three times per iteration), exactly as the reference's module calls do.
"""
import os

import torch
import torch.nn as nn

from .. import graphs
from .. import ops as K

config = {'batch_size': 128, 'image_channel': 1, 'z_dim': 100, 'g_hidden': 64, 'd_hidden': 64, 'x_dim': 64,
          'epochs': 20, 'real_label': 1., 'fake_label': 0., 'lr': 2e-4, 'seed': 1}

G_CH = [100, 512, 256, 128, 64, 1]
G_HW = [1, 4, 8, 16, 32, 64]
D_CH = [1, 64, 128, 256, 512, 1]
D_HW = [64, 32, 16, 8, 4, 1]


def weights_init(m):
    """mnist_dcgan.py:63-69."""
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find('BatchNorm') != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class Generator(nn.Module):
    def __init__(self):
        super().__init__()
        layers = []
        for i in range(5):
            stride, pad = (1, 0) if i == 0 else (2, 1)
            layers.append(nn.ConvTranspose2d(G_CH[i], G_CH[i + 1], 4, stride, pad, bias=False))
            layers += [nn.BatchNorm2d(G_CH[i + 1]), nn.ReLU(True)] if i < 4 else [nn.Tanh()]
        self.main = nn.Sequential(*layers)

    def forward(self, input):
        """[B,100,1,1] -> [B,1,64,64]; train mode uses batch statistics (and updates the running buffers)."""
        plan = _forward_plan(self, input.shape[0], "G")
        return plan.g_forward(input, self.training)


class Discriminator(nn.Module):
    def __init__(self):
        super().__init__()
        layers = [nn.Conv2d(1, 64, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True)]
        for i in range(1, 4):
            layers += [nn.Conv2d(D_CH[i], D_CH[i + 1], 4, 2, 1, bias=False), nn.BatchNorm2d(D_CH[i + 1]),
                       nn.LeakyReLU(0.2, inplace=True)]
        layers += [nn.Conv2d(512, 1, 4, 1, 0, bias=False), nn.Sigmoid()]
        self.main = nn.Sequential(*layers)

    def forward(self, input):
        plan = _forward_plan(self, input.shape[0], "D")
        return plan.d_forward(input, self.training)


def _forward_plan(module, batch, which):
    cache = module.__dict__.setdefault("_pcg_plans", {})
    p = cache.get(batch)
    if p is not None and not (p.G if which == "G" else p.D).aliases(module):
        p = None                        # a training plan re-adopted the module since: this plan's arena is stale
    if p is None:
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("pcg_b200: modules must live on a CUDA device (there is no CPU fallback)")
        p = DcganPlan(batch, dev, use_graph=False)
        (p.adopt_g if which == "G" else p.adopt_d)(module)
        cache.clear()
        cache[batch] = p
    p.refresh()
    return p


class _ConvT:
    """One ConvTranspose2d layer of G expressed through the mirrored convolution's geometry."""

    def __init__(self, i, B):
        self.cin_t, self.cout_t = G_CH[i], G_CH[i + 1]
        self.hin, self.hout = G_HW[i], G_HW[i + 1]
        self.stride, self.pad = (1, 0) if i == 0 else (2, 1)
        # mirrored conv: input [B, hout, hout, cout_t] -> output [B, hin, hin, cin_t]
        self.geom = (B, self.hout, self.hout, self.cout_t, self.cin_t, 4, self.stride, self.pad)


class DcganPlan:
    def __init__(self, batch, device, lr=2e-4, betas=(0.5, 0.999), use_graph=True, tensor_cores=None, share=None,
                 operand_terms=None):
        """share: another DcganPlan (any batch size) whose parameter arenas, Adam state, BatchNorm buffers and packed
        weights this plan uses too — the reference keeps ONE optimizer state for the whole run (mnist_dcgan.py:133-134),
        so the tail batch of an epoch (60000 % 128 = 96) must not get fresh moments.
        tensor_cores: the 64..512-channel convolutions (12 of the 15 conv passes' FLOPs) on the tcgen05 kernels with
        bf16 operands; None = follow PCG_PRECISION (default bf16 -> on), False = exact fp32 on the CUDA cores."""
        self.B, self.lr, self.betas = batch, lr, betas
        self.tc = (os.environ.get("PCG_PRECISION", "bf16") != "fp32") if tensor_cores is None else bool(tensor_cores)
        # operand precision of the tensor-core convolutions: 1 = plain bf16 operands, fp32 accumulation and fp32 storage
        # (default: over 200 iterations its loss curves leave the fp32 oracle's no faster than the fp32 CUDA-core plan's
        # own do, profiles/exp_dcgan_precision_r2.md), 3 = bf16x3 (fp32-equivalent products, 3x the tensor work: the mode
        # the per-step parity tests pin to 1e-2)
        self.terms = int(os.environ.get("PCG_TC_TERMS", "1")) if operand_terms is None else int(operand_terms)
        # data parallel: one process per GPU, per-replica BatchNorm statistics, gradients summed by an
        # all-reduce before each optimizer step (D's before the G phase, which sees the updated D), 1/world folded
        # into Adam
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist is not None else 1
        dev = self.dev = torch.device(device)
        B = batch
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        # ---- parameters (torch layouts), flat arenas
        gnames, dnames = [], []
        for i in range(5):
            gnames.append((f"main.{3 * i}.weight", (G_CH[i], G_CH[i + 1], 4, 4)))
            if i < 4:
                gnames += [(f"main.{3 * i + 1}.weight", (G_CH[i + 1],)), (f"main.{3 * i + 1}.bias", (G_CH[i + 1],))]
        dnames.append(("main.0.weight", (64, 1, 4, 4)))
        idx = 2
        self.d_conv_names, self.d_bn_names = ["main.0"], [None]
        for i in range(1, 4):
            dnames += [(f"main.{idx}.weight", (D_CH[i + 1], D_CH[i], 4, 4)), (f"main.{idx + 1}.weight", (D_CH[i + 1],)),
                       (f"main.{idx + 1}.bias", (D_CH[i + 1],))]
            self.d_conv_names.append(f"main.{idx}")
            self.d_bn_names.append(f"main.{idx + 1}")
            idx += 3
        dnames.append(("main.11.weight", (1, 512, 4, 4)))
        self.d_conv_names.append("main.11")
        if share is not None:
            self.G, self.D, self.D_grad2 = share.G, share.D, share.D_grad2
        else:
            self.G, self.D = K.FlatParams(gnames, dev), K.FlatParams(dnames, dev)
            self.D_grad2 = torch.zeros_like(self.D.grad)
        # BN buffers: running_mean / running_var / num_batches_tracked
        self.g_bn = [dict(rm=z(G_CH[i + 1]), rv=torch.ones(G_CH[i + 1], device=dev),
                          nbt=torch.zeros((), dtype=torch.int64, device=dev), st=K.BNState(G_CH[i + 1], dev))
                     for i in range(4)]
        self.d_bn = [None] + [dict(rm=z(D_CH[i + 1]), rv=torch.ones(D_CH[i + 1], device=dev),
                                   nbt=torch.zeros((), dtype=torch.int64, device=dev),
                                   st=[K.BNState(D_CH[i + 1], dev) for _ in range(2)]) for i in range(1, 4)]
        # ---- packed weights
        self.convt = [_ConvT(i, B) for i in range(5)]
        if share is not None:
            self.g_bn, self.d_bn = share.g_bn, share.d_bn           # batch-size independent (per-channel vectors)
            self.g_wf, self.g_wd, self.d_wf, self.d_wd = share.g_wf, share.g_wd, share.d_wf, share.d_wd
        else:
            self.g_wf = [z(G_CH[i] * G_CH[i + 1] * 16) for i in range(5)]
            self.g_wd = [z(G_CH[i] * G_CH[i + 1] * 16) for i in range(5)]
            self.d_wf = [z(D_CH[i] * D_CH[i + 1] * 16) for i in range(5)]
            self.d_wd = [z(D_CH[i] * D_CH[i + 1] * 16) for i in range(5)]
        # ---- static inputs / activations (NHWC)
        self.real, self.noise = z(B, 64, 64, 1), z(B, 1, 1, 100)
        self.gy = [z(B, G_HW[i + 1], G_HW[i + 1], G_CH[i + 1]) for i in range(5)]       # ConvT outputs (pre-BN)
        self.ga = [z(B, G_HW[i + 1], G_HW[i + 1], G_CH[i + 1]) for i in range(5)]       # post BN+ReLU / tanh
        self.gd = [z(B, G_HW[i + 1], G_HW[i + 1], G_CH[i + 1]) for i in range(5)]       # gradients wrt ga
        self.gdy = [z(B, G_HW[i + 1], G_HW[i + 1], G_CH[i + 1]) for i in range(5)]      # gradients wrt gy
        # discriminator: two activation sets (real pass, fake pass; the G-step pass reuses set 1)
        mk = lambda: [z(B, D_HW[i + 1], D_HW[i + 1], D_CH[i + 1]) for i in range(5)]  # noqa: E731
        self.dy = [mk(), mk()]
        self.da = [mk(), mk()]
        self.dd = mk()       # gradient wrt da (scratch, one pass at a time)
        self.ddy = mk()      # gradient wrt dy
        self.dz = z(B)
        self.dfake = z(B, 64, 64, 1)
        self.scal = z(8)     # 0 errD, 1 errG, 2 errD_real, 3 errD_fake, 4 D_x, 5 D_G_z1, 6 D_G_z2
        wmax = 0
        for i in range(5):
            wmax = max(wmax, int(K.conv_wgrad_scratch(*self.convt[i].geom, dev).numel()),
                       int(K.conv_wgrad_scratch(B, D_HW[i], D_HW[i], D_CH[i], D_CH[i + 1], 4, *self._dsp(i), dev).numel()))
        self.wsc = z(wmax)
        self.use_graph, self.graph = use_graph, None
        self.refresh()

    @staticmethod
    def _dsp(i):
        return (1, 0) if i == 4 else (2, 1)

    # ------------------------------------------------------------------ binding
    def adopt_g(self, module):
        self.G.adopt(module)
        bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm2d)]
        for b, m in zip(self.g_bn, bns):
            for key, name in (("rm", "running_mean"), ("rv", "running_var"), ("nbt", "num_batches_tracked")):
                b[key].copy_(getattr(m, name))
                m._buffers[name] = b[key]
        self.refresh()

    def adopt_d(self, module):
        self.D.adopt(module)
        bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm2d)]
        for b, m in zip(self.d_bn[1:], bns):
            for key, name in (("rm", "running_mean"), ("rv", "running_var"), ("nbt", "num_batches_tracked")):
                b[key].copy_(getattr(m, name))
                m._buffers[name] = b[key]
        self.refresh()

    def refresh(self):
        self._pack_g()
        self._pack_d()

    def _pack_g(self):
        for i in range(5):
            # ConvT weight [Cin_T][Cout_T][4][4] == mirrored conv weight OIHW with O = Cin_T, I = Cout_T
            K.pack_weights(self.G.p(f"main.{3 * i}.weight"), 4, wf=self.g_wf[i], wd=self.g_wd[i])

    def _pack_d(self):
        for i in range(5):
            K.pack_weights(self.D.p(self.d_conv_names[i] + ".weight"), 4, wf=self.d_wf[i], wd=self.d_wd[i])

    # ------------------------------------------------------------------ generator
    def _g_fwd(self, training=True):
        x = self.noise
        for i, L in enumerate(self.convt):
            K.conv_dgrad(x, *L.geom[:4], self.g_wd[i], *L.geom[4:], self.gy[i])          # ConvTranspose2d forward
            if i < 4:
                bn, nm = self.g_bn[i], f"main.{3 * i + 1}"
                C, M = G_CH[i + 1], self.B * G_HW[i + 1] ** 2
                if training:
                    K.bn_train_fwd(self.gy[i], M, C, self.G.p(nm + ".weight"), self.G.p(nm + ".bias"), bn["rm"], bn["rv"],
                                   bn["nbt"], bn["st"], self.ga[i], act=K.ACT_RELU)
                else:
                    K.bn_eval(self.gy[i], self.G.p(nm + ".weight"), self.G.p(nm + ".bias"), bn["rm"], bn["rv"], self.ga[i])
                    K.unary(self.ga[i], K.RELU, self.ga[i])
            else:
                K.unary(self.gy[i], K.TANH, self.ga[i])
            x = self.ga[i]

    def _g_bwd(self):
        """dfake (gradient wrt the tanh output) -> generator parameter gradients."""
        K.unary_bwd(self.dfake, self.ga[4], K.TANH, self.gdy[4])
        for i in range(4, -1, -1):
            L = self.convt[i]
            xin = self.noise if i == 0 else self.ga[i - 1]
            K.conv_wgrad(self.gdy[i], xin, *L.geom, self.wsc, self.G.g(f"main.{3 * i}.weight"))
            if i == 0:
                break
            K.conv_fprop(self.gdy[i], *L.geom[:4], self.g_wf[i], *L.geom[4:], self.gd[i - 1])   # ConvT input gradient
            nm, bn = f"main.{3 * (i - 1) + 1}", self.g_bn[i - 1]
            K.bn_train_bwd(self.gd[i - 1], self.gy[i - 1], self.B * G_HW[i] ** 2, G_CH[i], self.G.p(nm + ".weight"),
                           bn["st"], self.gdy[i - 1], self.G.g(nm + ".weight"), self.G.g(nm + ".bias"), act=K.ACT_RELU)

    # ------------------------------------------------------------------ discriminator
    def _d_fwd(self, x, s, training=True):
        """Forward on NHWC image batch x into activation set s; logits land in dy[s][4]."""
        for i in range(5):
            st, pd = self._dsp(i)
            H, Ci, Co = D_HW[i], D_CH[i], D_CH[i + 1]
            if i == 0:
                K.conv_fprop(x, self.B, H, H, Ci, self.d_wf[0], Co, 4, st, pd, self.da[s][0], act=K.ACT_LRELU, slope=0.2)
            else:
                K.conv_fprop(x, self.B, H, H, Ci, self.d_wf[i], Co, 4, st, pd, self.dy[s][i])
                if i < 4:
                    nm, bn = self.d_bn_names[i], self.d_bn[i]
                    if training:
                        K.bn_train_fwd(self.dy[s][i], self.B * D_HW[i + 1] ** 2, Co, self.D.p(nm + ".weight"),
                                       self.D.p(nm + ".bias"), bn["rm"], bn["rv"], bn["nbt"], bn["st"][s], self.da[s][i],
                                       act=K.ACT_LRELU, slope=0.2)
                    else:
                        K.bn_eval(self.dy[s][i], self.D.p(nm + ".weight"), self.D.p(nm + ".bias"), bn["rm"], bn["rv"],
                                  self.da[s][i])
                        K.unary(self.da[s][i], K.LRELU, self.da[s][i], 0.2)
            x = self.da[s][i] if i < 4 else None

    def _d_bwd(self, x, s, grad, want_wgrad, want_dx):
        """Backward from self.dz (gradient wrt the pre-sigmoid logit) through activation set s.
        ``grad(name)`` returns the gradient slot of a parameter; dx (wrt the image) goes to self.dfake."""
        B = self.B
        dout = self.dz        # [B] == [B,1,1,1]
        for i in range(4, -1, -1):
            st, pd = self._dsp(i)
            H, Ci, Co = D_HW[i], D_CH[i], D_CH[i + 1]
            xin = x if i == 0 else self.da[s][i - 1]
            if want_wgrad:
                K.conv_wgrad(xin, dout, B, H, H, Ci, Co, 4, st, pd, self.wsc, grad(self.d_conv_names[i] + ".weight"))
            if i == 0:
                if want_dx:
                    K.conv_dgrad(dout, B, H, H, Ci, self.d_wd[0], Co, 4, st, pd, self.dfake)
                break
            if i - 1 == 0:
                # layer 0 has no BN: multiply by LeakyReLU'(a0) in the dgrad epilogue
                K.conv_dgrad(dout, B, H, H, Ci, self.d_wd[i], Co, 4, st, pd, self.ddy[0], act_ref=self.da[s][0],
                             ref_act=K.ACT_LRELU, ref_slope=0.2)
            else:
                K.conv_dgrad(dout, B, H, H, Ci, self.d_wd[i], Co, 4, st, pd, self.dd[i - 1])
                nm, bn = self.d_bn_names[i - 1], self.d_bn[i - 1]
                K.bn_train_bwd(self.dd[i - 1], self.dy[s][i - 1], B * D_HW[i] ** 2, Ci, self.D.p(nm + ".weight"),
                               bn["st"][s], self.ddy[i - 1], grad(nm + ".weight"), grad(nm + ".bias"), act=K.ACT_LRELU,
                               slope=0.2)
            dout = self.ddy[i - 1]

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

    def _body(self):
        def both():
            self._d_phase()
            self._g_phase()
        self._tc(both)

    def _segments(self):
        """The iteration cut at the two gradient reductions: [D grads] all-reduce [D update, G grads] all-reduce
        [G update]."""
        def mid():
            self._d_update()
            self._g_grads()
        return [lambda: self._tc(self._d_grads), lambda: self._tc(mid), lambda: self._tc(self._g_update)]

    def _d_phase(self):
        self._d_grads()
        self._d_update()

    def _g_phase(self):
        self._g_grads()
        self._g_update()

    def _d_grads(self):
        B, D, G = self.B, self.D, self.G
        g1 = lambda n: D.g(n)  # noqa: E731
        g2 = lambda n: D._view(self.D_grad2, n)  # noqa: E731
        # (1) D on real (:147-153)
        self._d_fwd(self.real, 0)
        K.gan_loss(self.dy[0][4].view(-1), K.GAN_BCE, 1.0, self.scal[2:3], self.dz, out_aux=self.scal[4:5])
        self._d_bwd(self.real, 0, g1, True, False)
        # D on fake.detach() (:155-162)
        self._g_fwd()
        self._d_fwd(self.ga[4], 1)
        K.gan_loss(self.dy[1][4].view(-1), K.GAN_BCE, 0.0, self.scal[3:4], self.dz, out_aux=self.scal[5:6])
        self._d_bwd(self.ga[4], 1, g2, True, False)
        K.binary(D.grad, self.D_grad2, K.ADD, D.grad)                 # gradients accumulate over the two backward calls
        K.combine([(1.0, self.scal[2:3]), (1.0, self.scal[3:4])], self.scal[0:1])

    def _d_update(self):
        D = self.D
        b1, b2 = self.betas
        K.adam(D.data, D.grad, D.m, D.v, D.step, self.lr, b1, b2, grad_scale=1.0 / self.world)   # optimizerD.step() :164
        self._pack_d()

    def _g_grads(self):
        B, D, G = self.B, self.D, self.G
        g2 = lambda n: D._view(self.D_grad2, n)  # noqa: E731
        # (2) G through the updated D (:169-175)
        self._d_fwd(self.ga[4], 1)
        K.gan_loss(self.dy[1][4].view(-1), K.GAN_BCE, 1.0, self.scal[1:2], self.dz, out_aux=self.scal[6:7])
        self._d_bwd(self.ga[4], 1, g2, False, True)
        self._g_bwd()

    def _g_update(self):
        G = self.G
        b1, b2 = self.betas
        K.adam(G.data, G.grad, G.m, G.v, G.step, self.lr, b1, b2, grad_scale=1.0 / self.world)
        self._pack_g()

    def _state(self):
        t = [self.G.data, self.G.m, self.G.v, self.G.step, self.D.data, self.D.m, self.D.v, self.D.step]
        for b in self.g_bn + self.d_bn[1:]:
            t += [b["rm"], b["rv"], b["nbt"]]
        return t

    def step(self, real, noise):
        """real [B,1,64,64] (NCHW == NHWC for one channel), noise [B,100,1,1]; returns the scalar block."""
        self.real.view(-1).copy_(real.reshape(-1), non_blocking=True)
        self.noise.view(-1).copy_(noise.reshape(-1), non_blocking=True)
        if self.dist is not None:
            return self._step_dp()
        if not self.use_graph:
            self._body()
            return self.scal
        if self.graph is None:
            self._dry_run()
            self.graph = self._capture(self._body)
        self.graph.replay()
        return self.scal

    def _dry_run(self):
        """One eager iteration on a state snapshot (sizes the library scratch, sets function attributes)."""
        snap = [t.clone() for t in self._state()]
        self._body()
        torch.cuda.synchronize()
        for dst, src in zip(self._state(), snap):
            dst.copy_(src)
        self.refresh()
        torch.cuda.synchronize()

    @staticmethod
    def _capture(fn):
        return graphs.capture(fn)

    def _step_dp(self):
        segs = self._segments()
        if self.use_graph:
            if self.graph is None:
                self._dry_run()
                self.graph = [self._capture(f) for f in segs]
            segs = [g.replay for g in self.graph]
        segs[0]()
        self.dist.all_reduce(self.D.grad)
        segs[1]()
        self.dist.all_reduce(self.G.grad)
        segs[2]()
        return self.scal

    # ------------------------------------------------------------------ module forwards
    def g_forward(self, noise, training):
        with torch.no_grad():
            self.noise.view(-1).copy_(noise.reshape(-1))
            self._g_fwd(training)
            return self.ga[4].view(self.B, 1, 64, 64).clone()

    def d_forward(self, x, training):
        with torch.no_grad():
            self.real.view(-1).copy_(x.reshape(-1))
            self._d_fwd(self.real, 0, training)
            out = torch.empty(self.B, device=self.dev)
            K.unary(self.dy[0][4].view(-1), K.SIGMOID, out)
            return out


def train_dcgan(netG, netD, dataloader, cfg, device="cuda"):
    """The training loop of mnist_dcgan.py:140-193 as a function.  Returns (G_losses, D_losses) per epoch."""
    plans = {}            # one plan per batch size; all share the first plan's arenas, Adam state and BN buffers
    epoch_G, epoch_D = [], []
    for epoch in range(cfg['epochs']):
        acc = torch.zeros(8, device=device)
        n = 0
        for i, data in enumerate(dataloader):
            real = data[0].to(device, non_blocking=True).float()
            b = real.size(0)
            plan = plans.get(b)
            if plan is None:
                first = next(iter(plans.values()), None)
                plan = plans[b] = DcganPlan(b, device, cfg['lr'], share=first)
                if first is None:
                    plan.adopt_g(netG)
                    plan.adopt_d(netD)
            noise = torch.randn(b, cfg['z_dim'], 1, 1, device=device)
            sc = plan.step(real, noise)
            acc += sc
            n += 1
            if i % 200 == 0:
                v = sc.tolist()
                print(f"[{epoch}/{cfg['epochs']}][{i}/{len(dataloader)}] Loss_D: {v[0]:.4f} Loss_G: {v[1]:.4f} "
                      f"D(x): {v[4]:.4f} D(G(z)): {v[5]:.4f} / {v[6]:.4f}")
        tot = acc.tolist()
        epoch_G.append(tot[1] / max(n, 1))
        epoch_D.append(tot[0] / max(n, 1))
        print(f"Epoch [{epoch+1}/{cfg['epochs']}] Avg Loss_D: {epoch_D[-1]:.4f} Avg Loss_G: {epoch_G[-1]:.4f}")
    return epoch_G, epoch_D
