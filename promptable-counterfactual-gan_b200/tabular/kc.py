"""Native mirror of ``conditional_counteRGAN/house_sales_kc_usa``.

    models/generator.py:6-92       FiLM, ResidualBlock, ResidualGenerator(input_dim, hidden_dim, num_classes, continuous_idx,
                                   categorical_info, n_blocks=5, residual_scaling=0.1, tau=0.5)
                                   .forward(x, target_onehot, mask=None, temperature=None, hard=False)
    models/discriminator.py:5-20   Discriminator(input_dim, hidden_dim, num_classes)  (spectral-norm MLP critic)
    models/nn_classifier.py:4-32   NNClassifier(input_dim, output_dim=4)  (frozen, eval: BatchNorm running stats)
    trainer.py:186-378             train_countergan(generator, config, X_train, y_train, clf_model)

The Gumbel noise is an explicit input of the native step (drawn with torch on the device, exactly the
``-log(Exp(1))`` torch's ``F.gumbel_softmax`` uses), so the step itself is deterministic.
"""
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm

from .. import ops as K
from .layers import BN1d, Critic, Ctx, Dense, GraphStep


# ------------------------------------------------------------------ mirror modules
class FiLM(nn.Module):
    def __init__(self, hidden_dim, cond_dim):
        super().__init__()
        self.gamma = nn.Linear(cond_dim, hidden_dim)
        self.beta = nn.Linear(cond_dim, hidden_dim)


class ResidualBlock(nn.Module):
    def __init__(self, hidden_dim, cond_dim):
        super().__init__()
        self.fc1 = nn.Linear(hidden_dim, hidden_dim)
        self.bn1 = nn.BatchNorm1d(hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.bn2 = nn.BatchNorm1d(hidden_dim)
        self.film = FiLM(hidden_dim, cond_dim)


class ResidualGenerator(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_classes, continuous_idx, categorical_info, n_blocks=5,
                 residual_scaling=0.1, tau=0.5):
        super().__init__()
        self.input_dim, self.hidden_dim, self.num_classes = input_dim, hidden_dim, num_classes
        self.continuous_idx = list(continuous_idx)
        self.categorical_info = categorical_info
        self.cond_dim = input_dim + num_classes
        self.fc_in = nn.Linear(input_dim + self.cond_dim, hidden_dim)
        self.blocks = nn.ModuleList([ResidualBlock(hidden_dim, self.cond_dim) for _ in range(n_blocks)])
        self.fc_cont = nn.Linear(hidden_dim, len(self.continuous_idx))
        self.fc_cat_logits = nn.ModuleDict({str(idx): nn.Linear(hidden_dim, info["n"])
                                            for idx, info in self.categorical_info.items()})
        self.residual_scaling = residual_scaling
        self.tau = tau

    def forward(self, x, target_onehot, mask=None, temperature=None, hard=False):
        """(cont_residual, cat_logits, cat_samples) as generator.py:68-92.  ``hard=False`` is what the training path uses
        (trainer.py:259-261); ``hard=True`` (the evaluation path, eval_utils.py:75) returns the one-hot of the soft
        sample's arg-max, the forward value of F.gumbel_softmax(hard=True)."""
        if mask is None:
            mask = torch.ones_like(x)
        plan = _forward_plan(self, x.shape[0])
        tau = self.tau if temperature is None else float(temperature)
        noise = [torch.empty(x.shape[0], info["n"], device=x.device).exponential_() for info in self.categorical_info.values()]
        return plan.g_forward(x, target_onehot, mask, noise, tau, self.training, hard=hard)


class Discriminator(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_classes):
        super().__init__()
        h = hidden_dim
        self.net = nn.Sequential(
            spectral_norm(nn.Linear(input_dim + num_classes, h)), nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Linear(h, h * 2)), nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Linear(h * 2, h * 4)), nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Linear(h * 4, 1)))


class NNClassifier(nn.Module):
    def __init__(self, input_dim, output_dim=4):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(input_dim, 256), nn.LeakyReLU(0.1), nn.BatchNorm1d(256), nn.Dropout(0.3),
            nn.Linear(256, 256), nn.LeakyReLU(0.1), nn.BatchNorm1d(256), nn.Dropout(0.2),
            nn.Linear(256, 128), nn.LeakyReLU(0.1), nn.BatchNorm1d(128), nn.Dropout(0.1),
            nn.Linear(128, 64), nn.LeakyReLU(0.1), nn.BatchNorm1d(64),
            nn.Linear(64, output_dim))


def _forward_plan(gen, batch):
    cache = gen.__dict__.setdefault("_pcg_plans", {})
    p = cache.get(batch)
    if p is not None and not p.G.aliases(gen):
        p = None                        # a training plan re-adopted the module since: this plan's arena is stale
    if p is None:
        dev = next(gen.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("pcg_b200: modules must live on a CUDA device (there is no CPU fallback)")
        p = KcPlan(batch, dev, gen.categorical_info, gen.continuous_idx, input_dim=gen.input_dim, hidden=gen.hidden_dim,
                   nc=gen.num_classes, n_blocks=len(gen.blocks), tau=gen.tau, use_graph=False)
        p.adopt_g(gen)
        cache.clear()
        cache[batch] = p
    p.refresh()
    return p


# ------------------------------------------------------------------ plan
class KcPlan:
    def __init__(self, batch, device, categorical_info, continuous_idx, norm_vals=None, immutable_idx=(), input_dim=17,
                 hidden=32, nc=4, n_blocks=5, tau=0.5, lr_g=1e-3, lr_d=1e-3, lam=(2.0, 1.0, 1.0), use_graph=True):
        self.B, self.d, self.h, self.nc, self.nb, self.tau = batch, input_dim, hidden, nc, n_blocks, tau
        self.lr_g, self.lr_d, self.lam = lr_g, lr_d, lam
        self.cat = OrderedDict((int(k), int(v["n"])) for k, v in categorical_info.items())
        self.cont = list(continuous_idx)
        dev = self.dev = torch.device(device)
        ctx = self.ctx = Ctx(batch, dev)
        z, B, h, d = ctx.z, batch, hidden, input_dim
        cond = self.cond_dim = d + nc
        # ---- generator parameters (reference parameters() order)
        gn = [("fc_in.weight", (h, d + cond)), ("fc_in.bias", (h,))]
        for i in range(n_blocks):
            p = f"blocks.{i}."
            gn += [(p + "fc1.weight", (h, h)), (p + "fc1.bias", (h,)), (p + "bn1.weight", (h,)), (p + "bn1.bias", (h,)),
                   (p + "fc2.weight", (h, h)), (p + "fc2.bias", (h,)), (p + "bn2.weight", (h,)), (p + "bn2.bias", (h,)),
                   (p + "film.gamma.weight", (h, cond)), (p + "film.gamma.bias", (h,)),
                   (p + "film.beta.weight", (h, cond)), (p + "film.beta.bias", (h,))]
        gn += [("fc_cont.weight", (len(self.cont), h)), ("fc_cont.bias", (len(self.cont),))]
        for f, n in self.cat.items():
            gn += [(f"fc_cat_logits.{f}.weight", (n, h)), (f"fc_cat_logits.{f}.bias", (n,))]
        self.G = K.FlatParams(gn, dev)
        G = self.G
        self.fc_in = Dense(ctx, G, "fc_in", d + cond, h)
        self.blk = []
        for i in range(n_blocks):
            p = f"blocks.{i}."
            self.blk.append(dict(fc1=Dense(ctx, G, p + "fc1", h, h), bn1=BN1d(ctx, G, p + "bn1", h),
                                 fc2=Dense(ctx, G, p + "fc2", h, h), bn2=BN1d(ctx, G, p + "bn2", h),
                                 fg=Dense(ctx, G, p + "film.gamma", cond, h), fb=Dense(ctx, G, p + "film.beta", cond, h),
                                 g=z(B, h), b=z(B, h), u1=z(B, h), n1=z(B, h), f1=z(B, h), r1=z(B, h), u2=z(B, h), n2=z(B, h),
                                 f2=z(B, h), hin=None, hout=z(B, h), dg=z(B, h), db=z(B, h)))
        self.fc_cont = Dense(ctx, G, "fc_cont", h, len(self.cont))
        self.heads = OrderedDict((f, Dense(ctx, G, f"fc_cat_logits.{f}", h, n)) for f, n in self.cat.items())
        self.D = Critic(ctx, [(d + nc, h), (h, 2 * h), (2 * h, 4 * h), (4 * h, 1)], dev)
        # ---- classifier (frozen, eval)
        cdims = [(d, 256), (256, 256), (256, 128), (128, 64)]
        cn, idx, self.c_lin_names, self.c_bn_names = [], 0, [], []
        for j, (a, b) in enumerate(cdims):
            cn += [(f"net.{idx}.weight", (b, a)), (f"net.{idx}.bias", (b,)), (f"net.{idx + 2}.weight", (b,)),
                   (f"net.{idx + 2}.bias", (b,))]
            self.c_lin_names.append(f"net.{idx}")
            self.c_bn_names.append(f"net.{idx + 2}")
            idx += 4 if j < 3 else 3
        cn += [(f"net.{idx}.weight", (nc, 64)), (f"net.{idx}.bias", (nc,))]
        self.c_lin_names.append(f"net.{idx}")
        self.C = K.FlatParams(cn, dev)
        self.cl = [Dense(ctx, self.C, nm, a, b) for nm, (a, b) in zip(self.c_lin_names, cdims + [(64, nc)])]
        self.c_rm = [z(b) for _, b in cdims]
        self.c_rv = [torch.ones(b, device=dev) for _, b in cdims]
        self.c_scale = [z(b) for _, b in cdims]
        self.c_act = [z(B, b) for _, b in cdims]        # LeakyReLU(Linear) outputs
        self.c_bn = [z(B, b) for _, b in cdims]         # BatchNorm (eval) outputs
        self.c_d = [z(B, b) for _, b in cdims]
        self.c_d2 = [z(B, b) for _, b in cdims]
        self.clog, self.cdlog = z(B, nc), z(B, nc)
        self.clog0 = z(B, nc)
        # the training step sees the frozen classifier with every eval-mode BatchNorm folded into the Linear behind it
        # (refresh()): W' = W diag(s), b' = b + W t with s = gamma / sqrt(var + eps), t = beta - mean * s; five launches
        # forward and five backward (LeakyReLU' in the data-gradient epilogue) instead of nine and thirteen
        cfull = cdims + [(64, nc)]
        self.cwf = [None] + [z(b, a) for a, b in cfull[1:]]
        self.cwfT = [None] + [z(a, b) for a, b in cfull[1:]]
        self.cbf = [None] + [z(b) for _, b in cfull[1:]]
        self.c_dh = [z(B, b) for _, b in cdims]
        # ---- categorical value maps (trainer.py:205-224); default = the /(n-1) fallback
        self.nv = OrderedDict()
        for f, n in self.cat.items():
            v = norm_vals[f] if norm_vals is not None else torch.arange(n, dtype=torch.float32) / max(1.0, n - 1)
            self.nv[f] = v.to(dev).float().contiguous().view(1, n)
        # ---- static inputs
        self.x, self.mask = z(B, d), z(B, d)
        self.y_oh, self.t_oh = z(B, nc), z(B, nc)
        self.target = torch.zeros(B, dtype=torch.int64, device=dev)
        self.ones = torch.ones(B, d, device=dev)
        self.noise = OrderedDict((f, z(B, n)) for f, n in self.cat.items())     # Gumbel noise g = -log(Exp(1))
        # ---- activations
        self.cond, self.gin, self.h0 = z(B, cond), z(B, d + cond), z(B, h)
        self.contp, self.cont_out = z(B, len(self.cont)), z(B, len(self.cont))
        self.logits = OrderedDict((f, z(B, n)) for f, n in self.cat.items())
        self.samples = OrderedDict((f, z(B, n)) for f, n in self.cat.items())
        self.dsamples = OrderedDict((f, z(B, n)) for f, n in self.cat.items())
        self.dlogits = OrderedDict((f, z(B, n)) for f, n in self.cat.items())
        # per-head / per-block temporaries (not shared: a shared scratch tensor chains otherwise independent launches,
        # which the data-flow capture would have to serialise)
        self.svals = OrderedDict((f, z(B, 1)) for f in self.cat)
        self.dcols = OrderedDict((f, z(B, 1)) for f in self.cat)
        self.blk_tmp = [dict(dn2=z(B, h), du2=z(B, h), df=z(B, h), dn1=z(B, h), du1=z(B, h)) for _ in range(n_blocks)]
        self.dz_r, self.dz_f = z(B, 1), z(B, 1)
        self.res, self.masked, self.xcf, self.om, self.rm = (z(B, d) for _ in range(5))
        self.d_rm, self.d_pen, self.d_l1, self.d_masked, self.d_res, self.dx_adv, self.dx_cls = (z(B, d) for _ in range(7))
        self.d_contp = z(B, len(self.cont))
        self.dhA, self.dhB, self.tmp_h, self.df, self.dn, self.du = (z(B, h) for _ in range(6))
        self.dh_parts = [z(B, h) for _ in range(len(self.cat) + 1)]
        self.dz = z(B, 1)
        self.scal = z(16)   # 0 d_loss 1 g_loss 2 g_adv 3 g_cls 4 reg 5 mask_pen 6 pred_gain.. 7/8 d terms 9 D(real) 10 D(fake)
        # one launch per half residual block (csrc/film_layer.cu) where the shape allows; PCG_FILM_LAYER=0: the five / six
        # primitive operators
        import os
        self.fused = os.environ.get("PCG_FILM_LAYER", "1") != "0" and K.film_layer_supported(B, h)
        self.chain = os.environ.get("PCG_FILM_CHAIN", "1") != "0"        # forward: the ten half blocks as one chained call
        # the frozen classifier's forward + cross-entropy + input gradient as ONE launch (csrc/frozen_mlp.cu)
        self.cls_parts = K.frozen_mlp_parts([d] + [b for _, b in cdims] + [nc], B)
        self.fused_cls = os.environ.get("PCG_FROZEN_MLP", "1") != "0" and self.cls_parts > 0
        self.cls_part = z(max(self.cls_parts, 1))
        self.adv_part = z(max(self.cls_parts, 1))
        self.upd_part = [z(max(self.cls_parts, 1)) for _ in range(2)]
        # the critic's own update through the one-launch chain: wins at small batches (moons, 64 rows: 0.202 -> 0.179 ms),
        # loses at 4096 rows (0.657 -> 0.683 ms: the weight gradients no longer overlap the data-gradient chain)
        self.fused_update = os.environ.get("PCG_FUSED_CRITIC_UPDATE", "1" if B <= 1024 else "0") == "1"
        self.dzc = [torch.full((B, 1), -1.0 / B, device=dev), torch.full((B, 1), 1.0 / B, device=dev)]   # d(-+mean)/d score
        self.run = GraphStep(self._body, self._state, self.refresh, use_graph)
        self.refresh()

    # ---- binding
    def adopt_g(self, module):
        self.G.adopt(module)
        for blk, m in zip(self.blk, module.blocks):
            for bn, mb in ((blk["bn1"], m.bn1), (blk["bn2"], m.bn2)):
                bn.rm.copy_(mb.running_mean); bn.rv.copy_(mb.running_var); bn.nbt.copy_(mb.num_batches_tracked)
                mb._buffers["running_mean"], mb._buffers["running_var"], mb._buffers["num_batches_tracked"] = bn.rm, bn.rv, bn.nbt
        self.refresh()

    def adopt_d(self, module):
        self.D.adopt(module)

    def adopt_c(self, module):
        self.C.adopt(module)
        bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm1d)]
        for j, m in enumerate(bns):
            self.c_rm[j].copy_(m.running_mean)
            self.c_rv[j].copy_(m.running_var)
        self.refresh()

    def _all_dense(self):
        out = [self.fc_in, self.fc_cont] + list(self.heads.values()) + self.cl
        for b in self.blk:
            out += [b["fc1"], b["fc2"], b["fg"], b["fb"]]
        return out

    def refresh(self):
        K.transpose_multi([(L.W(), L.wT) for L in self._all_dense()])
        with torch.no_grad():
            for j in range(1, 5):
                nm = self.c_bn_names[j - 1]
                s = self.C.p(nm + ".weight") * torch.rsqrt(self.c_rv[j - 1] + 1e-5)
                t = self.C.p(nm + ".bias") - self.c_rm[j - 1] * s
                W = self.cl[j].W()
                torch.mul(W, s[None, :], out=self.cwf[j])
                self.cwfT[j].copy_(self.cwf[j].t())
                torch.addmv(self.cl[j].b(), W, t, out=self.cbf[j])

    def _state(self):
        t = [self.G.data, self.G.m, self.G.v, self.G.step, self.D.flat.data, self.D.flat.m, self.D.flat.v, self.D.flat.step]
        for b in self.blk:
            for bn in (b["bn1"], b["bn2"]):
                t += [bn.rm, bn.rv, bn.nbt]
        for L in self.D.layers:
            t += [L.u, L.v]
        return t

    # ---- generator forward (generator.py:68-92) + residual assembly (trainer.py:266-282)
    def _g_fwd(self, training=True):
        d, nc = self.d, self.nc
        K.copy_cols(self.t_oh, 0, self.cond, 0, nc)
        K.copy_cols(self.mask, 0, self.cond, nc, d)
        K.copy_cols(self.x, 0, self.gin, 0, d)
        K.copy_cols(self.cond, 0, self.gin, d, self.cond_dim)
        self.fc_in.fwd(self.gin, self.h0, K.ACT_RELU)
        hcur = self.h0
        if self.fused and training and self.chain:
            # the ten half blocks as ONE call of eleven launches; the FiLM tensors of every block first
            halves = []
            for b in self.blk:
                b["hin"] = hcur
                b["fg"].fwd(self.cond, b["g"])
                b["fb"].fwd(self.cond, b["b"])
                for fc, bn, u, n, out, res in ((b["fc1"], b["bn1"], b["u1"], b["n1"], b["r1"], None),
                                               (b["fc2"], b["bn2"], b["u2"], b["n2"], b["hout"], hcur)):
                    halves.append((fc.W(), fc.b(), bn.flat.p(bn.name + ".weight"), bn.flat.p(bn.name + ".bias"), bn.rm, bn.rv,
                                   bn.nbt, bn.st, b["g"], b["b"], res, u, n, out))
                hcur = b["hout"]
            K.film_chain_fwd(self.h0, halves)
            self.hlast = hcur
            self._g_heads(hcur)
            return
        for b in self.blk:
            b["hin"] = hcur
            b["fg"].fwd(self.cond, b["g"])
            b["fb"].fwd(self.cond, b["b"])
            if self.fused and training:
                for fc, bn, xin, u, n, out, res in ((b["fc1"], b["bn1"], hcur, b["u1"], b["n1"], b["r1"], None),
                                                    (b["fc2"], b["bn2"], b["r1"], b["u2"], b["n2"], b["hout"], hcur)):
                    K.film_layer_fwd(xin, fc.W(), fc.b(), bn.flat.p(bn.name + ".weight"), bn.flat.p(bn.name + ".bias"),
                                     bn.rm, bn.rv, bn.nbt, bn.st, b["g"], b["b"], u, n, out, res=res, relu=res is None)
                hcur = b["hout"]
                continue
            b["fc1"].fwd(hcur, b["u1"])
            b["bn1"].fwd(b["u1"], b["n1"], training=training)
            K.film_fwd(b["g"], b["n1"], b["b"], b["r1"], relu=True)                   # r1 = ReLU(FiLM(BN1(fc1 h)))
            b["fc2"].fwd(b["r1"], b["u2"])
            b["bn2"].fwd(b["u2"], b["n2"], training=training)
            K.film_fwd(b["g"], b["n2"], b["b"], b["hout"], res=hcur)                  # h' = h + FiLM(BN2(fc2 r1))
            hcur = b["hout"]
        self.hlast = hcur
        self._g_heads(hcur)

    def _g_heads(self, hcur):
        self.fc_cont.fwd(hcur, self.contp)
        K.unary(self.contp, K.SCALE, self.cont_out, 0.1)
        for f, head in self.heads.items():
            head.fwd(hcur, self.logits[f])
            K.gumbel_softmax_fwd(self.logits[f], self.noise[f], self.tau, self.samples[f])

    def _assemble(self):
        for i, f in enumerate(self.cont):
            K.copy_cols(self.cont_out, i, self.res, f, 1)
        for f in self.cat:
            K.linear_fwd(self.samples[f], self.nv[f], self.svals[f])             # onehot-like @ norm_vals
            K.copy_cols(self.svals[f], 0, self.res, f, 1)
            K.copy_cols(self.x, f, self.res, f, 1, alpha=-1.0, accumulate=True)  # - x[:, f]
        K.binary(self.res, self.mask, K.MUL, self.masked)
        K.binary(self.x, self.masked, K.ADD, self.xcf)

    def _c_fwd(self, x, logits):
        hcur = x
        for j in range(4):
            self.cl[j].fwd(hcur, self.c_act[j], K.ACT_LRELU, 0.1)
            nm = self.c_bn_names[j]
            K.bn_eval(self.c_act[j], self.C.p(nm + ".weight"), self.C.p(nm + ".bias"), self.c_rm[j], self.c_rv[j], self.c_bn[j],
                      scale_out=self.c_scale[j])
            hcur = self.c_bn[j]
        self.cl[4].fwd(hcur, logits)

    def _body(self):
        B, lam, d = self.B, self.lam, self.d
        D = self.D
        n = float(B * d)
        self._g_fwd()
        self._assemble()
        K.binary(self.ones, self.mask, K.ADD, self.om, 1.0, -1.0)
        K.binary(self.res, self.om, K.MUL, self.rm)
        K.reduce_scalar(self.rm, self.scal[5:6], 1.0 / n, absval=True, dx=self.d_rm, gscale=lam[2] / n)      # :287
        K.binary(self.d_rm, self.om, K.MUL, self.d_pen)
        # ---- D update (:290-295)
        if self.fused_cls and D.fused_ok and self.fused_update:
            # each pass: forward + backward chain as one launch, the weight gradients side by side; scalars off the path
            out_r = D.update_pass(self.x, self.y_oh, 0, -1.0, D.g1, self.upd_part[0], self.dzc[0])
            K.gan_loss(out_r, K.GAN_WASSERSTEIN, 1.0, self.scal[7:8], self.dz_r, out_aux=self.scal[9:10])
            out_f = D.update_pass(self.xcf, self.t_oh, 1, 1.0, D.g2, self.upd_part[1], self.dzc[1])
            K.gan_loss(out_f, K.GAN_WASSERSTEIN, 0.0, self.scal[8:9], self.dz_f)
        else:
            out_r = D.fwd(self.x, self.y_oh, 0)
            K.gan_loss(out_r, K.GAN_WASSERSTEIN, 1.0, self.scal[7:8], self.dz_r, out_aux=self.scal[9:10])
            D.bwd(self.dz_r, 0, D.g1)
            out_f = D.fwd(self.xcf, self.t_oh, 1)
            K.gan_loss(out_f, K.GAN_WASSERSTEIN, 0.0, self.scal[8:9], self.dz_f)
            D.bwd(self.dz_f, 1, D.g2)
        K.binary(D.flat.grad, D.grad2, K.ADD, D.flat.grad)
        K.combine([(1.0, self.scal[7:8]), (1.0, self.scal[8:9])], self.scal[0:1])
        D.flat.adam_step(self.lr_d)
        # ---- G update (:298-316)
        if self.fused_cls and D.fused_ok:
            # adversarial term: critic forward, -mean(score) and its input gradient as one launch; the scalars off the path
            out_g, ddin = D.input_gradient(self.xcf, self.t_oh, 1, -1.0, self.adv_part)
            K.gan_loss(out_g, K.GAN_WASSERSTEIN, 1.0, self.scal[2:3], self.dz, out_aux=self.scal[10:11])
        else:
            out_g = D.fwd(self.xcf, self.t_oh, 1)
            K.gan_loss(out_g, K.GAN_WASSERSTEIN, 1.0, self.scal[2:3], self.dz, out_aux=self.scal[10:11])
            ddin = D.bwd(self.dz, 1, None, want_dx=True)
        K.copy_cols(ddin, 0, self.dx_adv, 0, d)
        # frozen classifier, BatchNorm folded (see __init__): forward and input gradient
        if self.fused_cls:
            K.frozen_mlp_ce_grad([self.cl[0].W()] + self.cwf[1:], [self.cl[0].wT] + self.cwfT[1:],
                                 [self.cl[0].b()] + self.cbf[1:], self.xcf, self.target, self.cls_part, self.dx_cls,
                                 wgt=lam[0], slope=0.1, logits=self.clog)
            K.reduce_scalar(self.cls_part, self.scal[3:4], 1.0 / B)
        else:
            self._classifier_by_layers(lam)
        K.rownorm_mean(self.masked, 1, self.scal[4:5], dx=self.d_l1, gscale=lam[1])   # :305
        K.combine([(1.0, self.scal[2:3]), (lam[0], self.scal[3:4]), (lam[1], self.scal[4:5]), (lam[2], self.scal[5:6])],
                  self.scal[1:2])
        K.binary(self.dx_adv, self.dx_cls, K.ADD, self.d_masked)
        K.binary(self.d_masked, self.d_l1, K.ADD, self.d_masked)
        K.binary(self.d_masked, self.mask, K.MUL, self.d_res)
        K.binary(self.d_res, self.d_pen, K.ADD, self.d_res)
        self._g_bwd()
        self.G.adam_step(self.lr_g)
        g_layers = self._all_dense()[:2 + len(self.heads)]
        for b in self.blk:
            g_layers += [b[k] for k in ("fc1", "fc2", "fg", "fb")]
        K.transpose_multi([(L.W(), L.wT) for L in g_layers])            # dgrad operands of the updated generator
        # ---- diagnostics of trainer.py:319-343 that need an extra classifier pass are left to the caller

    def _classifier_by_layers(self, lam):
        self.cl[0].fwd(self.xcf, self.c_act[0], K.ACT_LRELU, 0.1)
        for j in range(1, 4):
            K.linear_fwd(self.c_act[j - 1], self.cwf[j], self.c_act[j], self.cbf[j], K.ACT_LRELU, 0.1)
        K.linear_fwd(self.c_act[3], self.cwf[4], self.clog, self.cbf[4])
        K.ce_loss(self.clog, self.target, self.scal[3:4], self.cdlog, wgt=lam[0])
        dcur = self.cdlog
        for j in range(4, 0, -1):
            K.linear_dgrad(dcur, self.cwfT[j], self.c_dh[j - 1], self.cwf[j].shape[1], act_ref=self.c_act[j - 1],
                           ref_act=K.ACT_LRELU, ref_slope=0.1)
            dcur = self.c_dh[j - 1]
        self.cl[0].dgrad(dcur, self.dx_cls)

    def _g_bwd(self):
        h = self.hlast
        # heads: continuous columns and the seven Gumbel-softmax heads feed d(res); dh accumulates in dhA
        for i, f in enumerate(self.cont):
            K.copy_cols(self.d_res, f, self.d_contp, i, 1, alpha=0.1)                 # * residual_scaling
        self.fc_cont.wgrad(h, self.d_contp)
        parts = self.dh_parts                        # one buffer per head: the eight data gradients run side by side ...
        self.fc_cont.dgrad(self.d_contp, parts[0])
        for i, (f, head) in enumerate(self.heads.items()):
            K.copy_cols(self.d_res, f, self.dcols[f], 0, 1)
            K.linear_dgrad(self.dcols[f], self.nv[f].view(-1, 1), self.dsamples[f], self.cat[f])   # d samples = d scalar * norm_vals
            K.softmax_bwd(self.dsamples[f], self.samples[f], self.tau, self.dlogits[f])
            head.wgrad(h, self.dlogits[f])
            head.dgrad(self.dlogits[f], parts[i + 1])
        live = list(parts[:len(self.heads) + 1])     # ... and meet in a tree of adds (a chain through add_src was 8 deep)
        while len(live) > 1:
            nxt = []
            for a in range(0, len(live) - 1, 2):
                dst = self.dhA if len(live) <= 2 else live[a]
                K.binary(live[a], live[a + 1], K.ADD, dst)
                nxt.append(dst)
            if len(live) % 2:
                nxt.append(live[-1])
            live = nxt
        if live[0] is not self.dhA:
            K.unary(live[0], K.COPY, self.dhA)
        dh, other = self.dhA, self.dhB
        for b, t in zip(reversed(self.blk), reversed(self.blk_tmp)):
            if self.fused:
                bn1, bn2, G = b["bn1"], b["bn2"], self.G
                K.film_layer_bwd(dh, b["g"], b["n2"], b["u2"], bn2.st, G.p(bn2.name + ".weight"), b["fc2"].W(), b["dg"],
                                 b["db"], t["du2"], t["df"], G.g(bn2.name + ".weight"), G.g(bn2.name + ".bias"),
                                 act_ref=b["r1"])                                     # d f1 (through the ReLU)
                b["fc2"].wgrad(b["r1"], t["du2"])
                K.film_layer_bwd(t["df"], b["g"], b["n1"], b["u1"], bn1.st, G.p(bn1.name + ".weight"), b["fc1"].W(), b["dg"],
                                 b["db"], t["du1"], other, G.g(bn1.name + ".weight"), G.g(bn1.name + ".bias"),
                                 add_src=dh, accumulate=True)                         # skip connection
                b["fc1"].wgrad(b["hin"], t["du1"])
                b["fg"].wgrad(self.cond, b["dg"])
                b["fb"].wgrad(self.cond, b["db"])
                dh, other = other, dh
                continue
            # f2 = g*n2 + b ; h' = h + f2
            K.film_bwd(dh, b["g"], b["n2"], t["dn2"], b["dg"], b["db"])              # d n2, d g / d b (first use)
            b["bn2"].bwd(t["dn2"], b["u2"], t["du2"])
            b["fc2"].wgrad(b["r1"], t["du2"])
            b["fc2"].dgrad(t["du2"], t["df"], act_ref=b["r1"], ref_act=K.ACT_RELU)    # d f1 (through the ReLU)
            K.film_bwd(t["df"], b["g"], b["n1"], t["dn1"], b["dg"], b["db"], accumulate=True)   # d n1, d g / d b (+=)
            b["bn1"].bwd(t["dn1"], b["u1"], t["du1"])
            b["fc1"].wgrad(b["hin"], t["du1"])
            b["fc1"].dgrad(t["du1"], other, add_src=dh)          # skip connection
            b["fg"].wgrad(self.cond, b["dg"])
            b["fb"].wgrad(self.cond, b["db"])
            dh, other = other, dh
        K.unary_bwd(dh, self.h0, K.RELU, dh)
        self.fc_in.wgrad(self.gin, dh)

    # ---- API
    def step(self, x, y, target, mask, exp_noise):
        """exp_noise: list of Exp(1) samples [B, n_f] in categorical_info order (what F.gumbel_softmax draws)."""
        self.x.copy_(x, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        K.onehot(y.contiguous(), self.nc, self.y_oh)
        K.onehot(self.target, self.nc, self.t_oh)
        for f, e in zip(self.cat, exp_noise):
            torch.neg(torch.log(e), out=self.noise[f])
        self.run()
        return self.scal

    def g_forward(self, x, target_onehot, mask, exp_noise, tau, training, hard=False):
        with torch.no_grad():
            self.x.copy_(x.float())
            self.t_oh.copy_(target_onehot.float())
            self.mask.copy_(mask.float())
            for f, e in zip(self.cat, exp_noise):
                torch.neg(torch.log(e), out=self.noise[f])
            old, self.tau = self.tau, tau
            self._g_fwd(training)
            self.tau = old
            if hard:
                for f in self.cat:
                    K.onehot_argmax(self.samples[f], self.samples[f])
            return (self.cont_out.clone(), {f: v.clone() for f, v in self.logits.items()},
                    {f: v.clone() for f, v in self.samples.items()})

    def diagnostics(self):
        """pred_gain / sparsity / l2 / class-flip of trainer.py:319-343 for the last step (one extra classifier pass)."""
        self._c_fwd(self.x, self.clog0)
        idx = torch.arange(self.B, device=self.dev)
        p_o = torch.softmax(self.clog0, 1)[idx, self.target]
        p_c = torch.softmax(self.clog, 1)[idx, self.target]
        return {"pred_gain": (p_c - p_o).mean().item(), "sparsity": 1.0 - (self.masked.abs() > 1e-3).float().mean().item(),
                "l2": self.masked.norm(dim=1).mean().item(), "flip": (self.clog.argmax(1) == self.target).float().mean().item()}


def train_countergan(generator, config, X_train, y_train, clf_model):
    """Drop-in for house_sales_kc_usa/trainer.py:186-378."""
    device = config['cuda']
    if not str(device).startswith("cuda"):
        raise RuntimeError("pcg_b200 needs a CUDA device (there is no CPU fallback)")
    torch.manual_seed(config.get('seed', 42))
    np.random.seed(config.get('seed', 42))
    num_classes = int(np.unique(y_train).size)
    bs = int(config.get('batch_size', 128))
    X_t = torch.tensor(X_train, dtype=torch.float32)
    y_t = torch.tensor(y_train, dtype=torch.long)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(X_t, y_t), batch_size=bs, shuffle=True, drop_last=True)
    cat, immutable = config['categorical_info'], config.get('immutable_idx', [])
    scaler, norm_vals = config.get('scaler', None), None
    if scaler is not None:                                       # trainer.py:205-216
        dmin = np.array(scaler.data_min_, dtype=float)
        drange = np.array(scaler.data_max_, dtype=float) - dmin
        norm_vals = {f: torch.tensor((np.array(i['raw_values'], dtype=float) - dmin[f]) / (drange[f] + 1e-12),
                                     dtype=torch.float32) for f, i in cat.items()}
    D = Discriminator(config['input_dim'], config['hidden_dim'], num_classes).to(device)
    plan = KcPlan(bs, device, cat, config['continuous_idx'], norm_vals, immutable, config['input_dim'], config['hidden_dim'],
                  num_classes, len(generator.blocks), config['gumbel_tau'], config['lr_G'], config['lr_D'],
                  (config['lambda_cls'], config['lambda_reg'], config['lambda_mask']))
    plan.adopt_g(generator.to(device))
    plan.adopt_d(D)
    plan.adopt_c(clf_model.to(device).eval())
    d_losses, g_losses = [], []
    for epoch in range(config['epochs']):
        acc, nb, diag = torch.zeros(16, device=device), 0, []
        for batch_idx, (xb, yb) in enumerate(loader):
            xb, yb = xb.to(device), yb.to(device)
            b, dd = xb.shape
            target_y = torch.randint(0, num_classes, (b,), device=device)
            target_y = torch.where(target_y == yb, (target_y + 1) % num_classes, target_y)
            mask = torch.randint(0, 2, (b, dd), device=device).float()
            if len(immutable) > 0:
                mask[:, immutable] = 0.0
            noise = [torch.empty(b, i["n"], device=device).exponential_() for i in cat.values()]
            sc = plan.step(xb, yb, target_y, mask, noise)
            acc += sc
            nb += 1
            if batch_idx % 100 == 0:
                v = sc.tolist()
                diag.append(plan.diagnostics())
                print(f"[Epoch {epoch+1}/{config['epochs']}] batch {batch_idx} :: D(real)={v[9]:.3f}, D(fake)={v[10]:.3f}, "
                      f"g_adv={v[2]:.4f}, g_cls={v[3]:.4f}, reg={v[4]:.6f}, mask_pen={v[5]:.5f}")
        tot = acc.tolist()
        d_losses.append(tot[0] / max(nb, 1))
        g_losses.append(tot[1] / max(nb, 1))
        m = {k: float(np.mean([q[k] for q in diag])) for k in diag[0]} if diag else {}
        print(f"[{epoch+1}/{config['epochs']}] D: {d_losses[-1]:.4f}, G: {g_losses[-1]:.4f}, " +
              ", ".join(f"{k}={v:.4f}" for k, v in m.items()))
    os.makedirs(config['out_dir'], exist_ok=True)
    torch.save(generator.state_dict(), config['generator_path'])
    print(f"Saved generator model to {config['generator_path']}")
    return d_losses, g_losses
