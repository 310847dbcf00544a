"""Native mirror of ``conditional_counteRGAN/moons``.

    models/generator.py:4-24      ResidualGenerator(input_dim, hidden_dim, num_classes).forward(x, target_onehot, mask)
    models/discriminator.py:6-22  Discriminator(input_dim, hidden_dim, num_classes).forward(x, target_onehot)
    models/nn_classifier.py:3-15  NNClassifier(input_dim, hidden_dim=32, num_classes=3).forward(x)
    trainer.py:31-128             train_countergan(generator, config, X_train, y_train, clf_model)
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from .. import ops as K
from .layers import BN1d, Critic, Ctx, Dense, GraphStep


# ------------------------------------------------------------------ mirror modules (parameter containers + native forward)
class ResidualGenerator(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_classes):
        super().__init__()
        self.dims = (input_dim, hidden_dim, num_classes)
        h = hidden_dim
        self.net = nn.Sequential(
            nn.Linear(input_dim + num_classes + input_dim, h), nn.BatchNorm1d(h), nn.ReLU(),
            nn.Linear(h, h), nn.BatchNorm1d(h), nn.ReLU(),
            nn.Linear(h, h // 2), nn.BatchNorm1d(h // 2), nn.ReLU(),
            nn.Linear(h // 2, input_dim))

    def forward(self, x, target_onehot, mask=None):
        plan = _forward_plan(self, x.shape[0])
        return plan.g_forward(x, target_onehot, mask, self.training)


class Discriminator(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_classes):
        super().__init__()
        h = hidden_dim
        self.net = nn.Sequential(
            spectral_norm(nn.Linear(input_dim + num_classes, h)), nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Linear(h, h // 2)), nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Linear(h // 2, h // 2)), nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Linear(h // 2, 1)))


class NNClassifier(nn.Module):
    def __init__(self, input_dim, hidden_dim=32, num_classes=3):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
                                 nn.Linear(hidden_dim, num_classes))


def _forward_plan(gen, batch):
    cache = gen.__dict__.setdefault("_pcg_plans", {})
    p = cache.get(batch)
    if p is not None and not p.G.aliases(gen):
        p = None                        # a training plan re-adopted the module since: this plan's arena is stale
    if p is None:
        dev = next(gen.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("pcg_b200: modules must live on a CUDA device (there is no CPU fallback)")
        i, h, nc = gen.dims
        p = MoonsPlan(batch, dev, i, h, nc, use_graph=False)
        p.adopt_g(gen)
        cache.clear()
        cache[batch] = p
    p.refresh()
    return p


# ------------------------------------------------------------------ plan
class MoonsPlan:
    def __init__(self, batch, device, input_dim=2, hidden=32, nc=3, lr_g=1e-3, lr_d=1e-3, lam=(2.0, 5.0, 5.0, 3.0),
                 use_graph=True):
        self.B, self.i, self.h, self.nc = batch, input_dim, hidden, nc
        self.lr_g, self.lr_d, self.lam = lr_g, lr_d, lam
        dev = self.dev = torch.device(device)
        ctx = self.ctx = Ctx(batch, dev)
        z, B, h = ctx.z, batch, hidden
        dims = [(2 * input_dim + nc, h), (h, h), (h, h // 2)]
        gn = []
        for j, (a, b) in enumerate(dims):
            gn += [(f"net.{3 * j}.weight", (b, a)), (f"net.{3 * j}.bias", (b,)), (f"net.{3 * j + 1}.weight", (b,)),
                   (f"net.{3 * j + 1}.bias", (b,))]
        gn += [("net.9.weight", (input_dim, h // 2)), ("net.9.bias", (input_dim,))]
        self.G = K.FlatParams(gn, dev)
        self.gl = [Dense(ctx, self.G, f"net.{3 * j}", a, b) for j, (a, b) in enumerate(dims)] + \
                  [Dense(ctx, self.G, "net.9", h // 2, input_dim)]
        self.gbn = [BN1d(ctx, self.G, f"net.{3 * j + 1}", b) for j, (_, b) in enumerate(dims)]
        self.D = Critic(ctx, [(input_dim + nc, h), (h, h // 2), (h // 2, h // 2), (h // 2, 1)], dev)
        cn = [("net.0.weight", (h, input_dim)), ("net.0.bias", (h,)), ("net.2.weight", (h, h)), ("net.2.bias", (h,)),
              ("net.4.weight", (nc, h)), ("net.4.bias", (nc,))]
        self.C = K.FlatParams(cn, dev)
        self.cl = [Dense(ctx, self.C, "net.0", input_dim, h), Dense(ctx, self.C, "net.2", h, h), Dense(ctx, self.C, "net.4", h, nc)]
        # static inputs
        self.x, self.mask = z(B, input_dim), z(B, input_dim)
        self.y_oh, self.t_oh = z(B, nc), z(B, nc)
        self.target = torch.zeros(B, dtype=torch.int64, device=dev)
        self.ones = torch.ones(B, input_dim, device=dev)
        # generator activations
        self.gin = z(B, 2 * input_dim + nc)
        self.gy = [z(B, b) for (_, b) in dims]
        self.ga = [z(B, b) for (_, b) in dims]
        self.gda = [z(B, b) for (_, b) in dims]
        self.gdy = [z(B, b) for (_, b) in dims]
        self.raw, self.masked, self.xcf, self.om, self.rm = (z(B, input_dim) for _ in range(5))
        self.d_rm, self.d_pen, self.d_l1, self.d_l2, self.d_masked, self.d_raw, self.dx_adv, self.dx_cls = (z(B, input_dim) for _ in range(8))
        self.dz = z(B, 1)
        self.dz_r, self.dz_f = z(B, 1), z(B, 1)
        # classifier activations
        self.c0, self.c1, self.clog, self.cdlog = z(B, h), z(B, h), z(B, nc), z(B, nc)
        self.cd1, self.cd0 = z(B, h), z(B, h)
        self.scal = z(16)   # 0 d_loss 1 g_loss 2 g_adv 3 g_cls 4 l1 5 l2 6 mask_pen 7 d_real_term 8 d_fake_term 9/10 sigmoid means
        import os
        cdims = [self.cl[0].k] + [L.n for L in self.cl]
        parts = K.frozen_mlp_parts(cdims, batch)
        self.fused_frozen = os.environ.get("PCG_FROZEN_MLP", "1") != "0" and parts > 0 and self.D.fused_ok
        self.cls_part, self.adv_part = ctx.z(max(parts, 1)), ctx.z(max(parts, 1))
        self.upd_part = [ctx.z(max(parts, 1)) for _ in range(2)]
        self.fused_update = os.environ.get("PCG_FUSED_CRITIC_UPDATE", "1" if batch <= 1024 else "0") == "1"   # see tabular/kc.py
        self.dzc = [torch.full((batch, 1), -1.0 / batch, device=dev), torch.full((batch, 1), 1.0 / batch, device=dev)]
        self.run = GraphStep(self._body, self._state, self.refresh, use_graph)
        self.refresh()

    # ---- binding
    def adopt_g(self, module):
        self.G.adopt(module)
        bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm1d)]
        for b, m in zip(self.gbn, bns):
            b.rm.copy_(m.running_mean); b.rv.copy_(m.running_var); b.nbt.copy_(m.num_batches_tracked)
            m._buffers["running_mean"], m._buffers["running_var"], m._buffers["num_batches_tracked"] = b.rm, b.rv, b.nbt
        self.refresh()

    def adopt_d(self, module):
        self.D.adopt(module)

    def adopt_c(self, module):
        self.C.adopt(module)
        self.refresh()

    def refresh(self):
        for L in self.gl + self.cl:
            L.refresh()

    def _state(self):
        t = [self.G.data, self.G.m, self.G.v, self.G.step, self.D.flat.data, self.D.flat.m, self.D.flat.v, self.D.flat.step]
        for b in self.gbn:
            t += [b.rm, b.rv, b.nbt]
        for L in self.D.layers:
            t += [L.u, L.v]
        return t

    # ---- pieces
    def _g_fwd(self, training=True):
        i, nc = self.i, self.nc
        K.copy_cols(self.x, 0, self.gin, 0, i)
        K.copy_cols(self.t_oh, 0, self.gin, i, nc)
        K.copy_cols(self.mask, 0, self.gin, i + nc, i)
        hcur = self.gin
        for j in range(3):
            self.gl[j].fwd(hcur, self.gy[j])
            self.gbn[j].fwd(self.gy[j], self.ga[j], act=K.ACT_RELU, training=training)
            hcur = self.ga[j]
        self.gl[3].fwd(hcur, self.raw)
        K.binary(self.raw, self.mask, K.MUL, self.masked)

    def _body(self):
        B, lam = self.B, self.lam
        D = self.D
        n = float(B * self.i)
        self._g_fwd()
        K.binary(self.x, self.masked, K.ADD, self.xcf)                        # x_cf = x + masked        trainer.py:69
        K.binary(self.ones, self.mask, K.ADD, self.om, 1.0, -1.0)              # 1 - mask
        K.binary(self.raw, self.om, K.MUL, self.rm)
        K.reduce_scalar(self.rm, self.scal[6:7], 1.0 / n, absval=True, dx=self.d_rm, gscale=lam[3] / n)   # :67
        K.binary(self.d_rm, self.om, K.MUL, self.d_pen)
        # ---- D update (:72-77)
        if self.fused_frozen and self.fused_update:
            out_r = D.update_pass(self.x, self.y_oh, 0, -1.0, D.g1, self.upd_part[0], self.dzc[0])
            K.gan_loss(out_r, K.GAN_WASSERSTEIN, 1.0, self.scal[7:8], self.dz_r, out_aux=self.scal[9:10])
            out_f = D.update_pass(self.xcf, self.t_oh, 1, 1.0, D.g2, self.upd_part[1], self.dzc[1])
            K.gan_loss(out_f, K.GAN_WASSERSTEIN, 0.0, self.scal[8:9], self.dz_f, out_aux=self.scal[10:11])
        else:
            out_r = D.fwd(self.x, self.y_oh, 0)
            K.gan_loss(out_r, K.GAN_WASSERSTEIN, 1.0, self.scal[7:8], self.dz_r, out_aux=self.scal[9:10])
            D.bwd(self.dz_r, 0, D.g1)
            out_f = D.fwd(self.xcf, self.t_oh, 1)
            K.gan_loss(out_f, K.GAN_WASSERSTEIN, 0.0, self.scal[8:9], self.dz_f, out_aux=self.scal[10:11])
            D.bwd(self.dz_f, 1, D.g2)
        K.binary(D.flat.grad, D.grad2, K.ADD, D.flat.grad)
        K.combine([(1.0, self.scal[7:8]), (1.0, self.scal[8:9])], self.scal[0:1])
        D.flat.adam_step(self.lr_d)
        # ---- G update (:80-95)
        if self.fused_frozen:
            # the critic's score and the frozen classifier's cross-entropy with their input gradients: one launch each
            out_g, ddin = D.input_gradient(self.xcf, self.t_oh, 1, -1.0, self.adv_part)
            K.gan_loss(out_g, K.GAN_WASSERSTEIN, 1.0, self.scal[2:3], self.dz)
            K.copy_cols(ddin, 0, self.dx_adv, 0, self.i)
            K.frozen_mlp_ce_grad([L.W() for L in self.cl], [L.wT for L in self.cl], [L.b() for L in self.cl], self.xcf,
                                 self.target, self.cls_part, self.dx_cls, wgt=lam[0], slope=0.0, logits=self.clog)
            K.reduce_scalar(self.cls_part, self.scal[3:4], 1.0 / B)
        else:
            out_g = D.fwd(self.xcf, self.t_oh, 1)
            K.gan_loss(out_g, K.GAN_WASSERSTEIN, 1.0, self.scal[2:3], self.dz)
            ddin = D.bwd(self.dz, 1, None, want_dx=True)
            K.copy_cols(ddin, 0, self.dx_adv, 0, self.i)
            self.cl[0].fwd(self.xcf, self.c0, K.ACT_RELU)
            self.cl[1].fwd(self.c0, self.c1, K.ACT_RELU)
            self.cl[2].fwd(self.c1, self.clog)
            K.ce_loss(self.clog, self.target, self.scal[3:4], self.cdlog, wgt=lam[0])
            self.cl[2].dgrad(self.cdlog, self.cd1, act_ref=self.c1, ref_act=K.ACT_RELU)
            self.cl[1].dgrad(self.cd1, self.cd0, act_ref=self.c0, ref_act=K.ACT_RELU)
            self.cl[0].dgrad(self.cd0, self.dx_cls)
        K.rownorm_mean(self.masked, 1, self.scal[4:5], dx=self.d_l1, gscale=lam[1])
        K.rownorm_mean(self.masked, 2, self.scal[5:6], dx=self.d_l2, gscale=lam[2])
        K.combine([(1.0, self.scal[2:3]), (lam[0], self.scal[3:4]), (lam[1], self.scal[4:5]), (lam[2], self.scal[5:6]),
                   (lam[3], self.scal[6:7])], self.scal[1:2])
        K.binary(self.dx_adv, self.dx_cls, K.ADD, self.d_masked)
        K.binary(self.d_masked, self.d_l1, K.ADD, self.d_masked)
        K.binary(self.d_masked, self.d_l2, K.ADD, self.d_masked)
        K.binary(self.d_masked, self.mask, K.MUL, self.d_raw)
        K.binary(self.d_raw, self.d_pen, K.ADD, self.d_raw)
        # generator backward
        self.gl[3].wgrad(self.ga[2], self.d_raw)
        self.gl[3].dgrad(self.d_raw, self.gda[2])
        for j in range(2, -1, -1):
            self.gbn[j].bwd(self.gda[j], self.gy[j], self.gdy[j], act=K.ACT_RELU)
            self.gl[j].wgrad(self.gin if j == 0 else self.ga[j - 1], self.gdy[j])
            if j > 0:
                self.gl[j].dgrad(self.gdy[j], self.gda[j - 1])
        self.G.adam_step(self.lr_g)
        for L in self.gl:
            L.refresh()

    def step(self, x, y, target, mask):
        self.x.copy_(x, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        K.onehot(y.contiguous(), self.nc, self.y_oh)
        K.onehot(self.target, self.nc, self.t_oh)
        self.run()
        return self.scal

    def g_forward(self, x, target_onehot, mask, training):
        with torch.no_grad():
            self.x.copy_(x.float())
            self.t_oh.copy_(target_onehot.float())
            self.mask.copy_(mask.float())
            self._g_fwd(training)
            return self.raw.clone(), self.masked.clone()


def train_countergan(generator, config, X_train, y_train, clf_model):
    """Drop-in for moons/trainer.py:31-128 (the critic is created here, as the reference does at :45)."""
    device = config['cuda']
    if not str(device).startswith("cuda"):
        raise RuntimeError("pcg_b200 needs a CUDA device (there is no CPU fallback)")
    seed = config['seed']
    torch.manual_seed(seed)
    np.random.seed(seed)
    num_classes = int(np.unique(y_train).size)
    X_t = torch.tensor(X_train, dtype=torch.float32)
    y_t = torch.tensor(y_train, dtype=torch.long)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(X_t, y_t), batch_size=config['batch_size'],
                                         shuffle=True, drop_last=True)
    D = Discriminator(config['input_dim'], config['hidden_dim'], num_classes).to(device)
    plan = MoonsPlan(config['batch_size'], device, config['input_dim'], config['hidden_dim'], num_classes, config['lr_G'],
                     config['lr_D'], (config['lambda_cls'], config['lambda_reg_l1'], config['lambda_reg_l2'],
                                      config['lambda_mask']))
    plan.adopt_g(generator.to(device))
    plan.adopt_d(D)
    plan.adopt_c(clf_model.to(device).eval())
    d_losses, g_losses = [], []
    for epoch in range(config['epochs']):
        acc, nb = torch.zeros(16, device=device), 0
        for batch_idx, (xb, yb) in enumerate(loader):
            xb, yb = xb.to(device), yb.to(device)
            bs, nf = xb.shape
            target_y = torch.randint(0, num_classes, (bs,), device=device)
            target_y = torch.where(target_y == yb, (target_y + 1) % num_classes, target_y)
            mask = torch.randint(0, 2, (bs, nf), device=device).float()
            sc = plan.step(xb, yb, target_y, mask)
            acc += sc
            nb += 1
            if (epoch + 1) % max(config['epochs'] * 0.1, 1) == 0 and batch_idx % 5 == 0:
                v = sc.tolist()
                print(f"[Epoch {epoch+1}/{config['epochs']}] batch {batch_idx} :: D(real)={v[9]:.3f}, D(fake)={v[10]:.3f}, "
                      f"g_adv={v[2]:.4f}, g_cls={v[3]:.4f}, reg_l1={v[4]:.5f}, reg_l2= {v[5]:.5f}, mask_pen={v[6]:.5f}")
        tot = acc.tolist()
        d_losses.append(tot[0] / max(nb, 1))
        g_losses.append(tot[1] / max(nb, 1))
        if (epoch + 1) % max(config['epochs'] * 0.2, 1) == 0:
            print(f"[{epoch+1}/{config['epochs']}] D: {d_losses[-1]:.4f}, G: {g_losses[-1]:.4f}")
    os.makedirs(config['out_dir'], exist_ok=True)
    torch.save(generator.state_dict(), config['generator_path'])
    print(f"Generator saved to {config['generator_path']}")
    return d_losses, g_losses
