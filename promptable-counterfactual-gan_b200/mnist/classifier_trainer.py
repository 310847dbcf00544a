"""Native mirror of classifier pre-training, ``conditional_counteRGAN/mnist/trainer.py:8-39``.

``train_classifier(classifier, train_loader, valid_loader, cfg, device)`` has the reference's signature and side effects
(per-epoch validation accuracy print, best state saved to ``cfg.classifier_path``).  One training iteration - the
CNNClassifier of models/classifier.py:4-28 in train mode (Dropout2d(0.25) after the conv stack, Dropout(0.5) after fc.1),
CrossEntropyLoss, backward, Adam(lr=cfg.cls_lr) - is composed from libpcg's primitive operators on NHWC fp32 tensors and
replayed as a data-flow captured CUDA graph (pcg_b200.dataflow); fc.1 runs as the 7x7 convolution it is on the NHWC map
(its torch weight [256, 128*7*7] IS the OIHW tensor [256, 128, 7, 7]), so no flatten permutation exists anywhere.
The dropout keep-masks are drawn on the device by ``pcg_dropout_mask`` (Philox; first nodes of the graph) or injected
(tests).  Validation uses the eval forward of the same plan.
"""
import os

import torch

from .. import dataflow
from .. import graphs
from .. import ops as K

CONVS = [("conv.0", 1, 32, 1, 28), ("conv.2", 32, 64, 2, 28), ("conv.4", 64, 128, 2, 14)]     # name, Cin, Cout, stride, H in


class ClassifierPlan:
    def __init__(self, batch, device, lr=1e-3, use_graph=True):
        self.B, self.lr = batch, lr
        dev = self.dev = torch.device(device)
        B = batch
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        names = []
        for nm, ci, co, _, _ in CONVS:
            names += [(nm + ".weight", (co, ci, 3, 3)), (nm + ".bias", (co,))]
        names += [("fc.1.weight", (256, 128 * 49)), ("fc.1.bias", (256,)), ("fc.4.weight", (10, 256)), ("fc.4.bias", (10,))]
        self.C = K.FlatParams(names, dev)
        # geometry of every layer as a convolution: (N, H, W, Cin, Cout, k, stride, pad)
        self.geom = [(B, h, h, ci, co, 3, st, 1) for _, ci, co, st, h in CONVS] + \
                    [(B, 7, 7, 128, 256, 7, 1, 0), (B, 1, 1, 256, 10, 1, 1, 0)]
        self.wname = [nm for nm, *_ in CONVS] + ["fc.1", "fc.4"]
        self.wf = [z(g[3] * g[4] * g[5] * g[5]) for g in self.geom]
        self.wd = [z(g[3] * g[4] * g[5] * g[5]) for g in self.geom]
        self.wsc = [K.conv_wgrad_scratch(*g, dev) for g in self.geom]
        self.stat = [K.stat_scratch(max(g[4], 4), dev) for g in self.geom]
        out_hw = [28, 14, 7, 1, 1]
        self.x = z(B, 28, 28, 1)
        self.y = torch.zeros(B, dtype=torch.int64, device=dev)
        self.a = [z(B, hw, hw, g[4]) for hw, g in zip(out_hw, self.geom)]      # ReLU outputs (a[4] = logits)
        self.d = [z(B, hw, hw, g[4]) for hw, g in zip(out_hw, self.geom)]      # gradients wrt the pre-activations
        self.a3d, self.da3d = z(B, 7, 7, 128), z(B, 7, 7, 128)
        self.f1d, self.df1d = z(B, 256), z(B, 256)
        self.m2, self.m1 = z(B, 7, 7, 128), z(B, 256)                          # dropout keep-masks (scaled)
        self.loss = z(1)
        self.rng = torch.zeros(3, dtype=torch.int64, device=dev)
        self.rng[2] = torch.initial_seed() & (2 ** 62 - 1)
        self.inject = False                         # True: m2 / m1 are written by the caller (tests)
        self.use_graph, self.graph = use_graph, None
        self.refresh()

    def adopt(self, module):
        self.C.adopt(module)
        self.refresh()

    def refresh(self):
        for i, nm in enumerate(self.wname):
            g = self.geom[i]
            w = self.C.p(nm + ".weight").view(g[4], g[3], g[5], g[5])
            K.pack_weights(w, g[5], wf=self.wf[i], wd=self.wd[i])

    # ------------------------------------------------------------------ forward
    def _fwd(self, train):
        h = self.x
        for i in range(3):
            K.conv_fprop(h, *self.geom[i][:4], self.wf[i], *self.geom[i][4:], self.a[i], bias=self.C.p(self.wname[i] + ".bias"),
                         act=K.ACT_RELU)
            h = self.a[i]
        if train:
            K.binary(h, self.m2, K.MUL, self.a3d)                              # Dropout2d(0.25), classifier.py:14
            h = self.a3d
        K.conv_fprop(h, *self.geom[3][:4], self.wf[3], *self.geom[3][4:], self.a[3], bias=self.C.p("fc.1.bias"), act=K.ACT_RELU)
        h = self.a[3].view(self.B, 256)
        if train:
            K.binary(h, self.m1, K.MUL, self.f1d)                              # Dropout(0.5), classifier.py:19
            h = self.f1d
        K.conv_fprop(h, *self.geom[4][:4], self.wf[4], *self.geom[4][4:], self.a[4], bias=self.C.p("fc.4.bias"))

    def _body(self):
        B = self.B
        if not self.inject:
            K.dropout_mask(self.m2, 0.25, channelwise=True, rng_state=self.rng)
            K.dropout_mask(self.m1, 0.5, rng_state=self.rng)
        self._fwd(True)
        logits, dlog = self.a[4].view(B, 10), self.d[4].view(B, 10)
        K.ce_loss(logits, self.y, self.loss, dlog)                             # CrossEntropyLoss (mean), trainer.py:19
        # fc.4
        K.conv_wgrad(self.f1d, dlog, *self.geom[4], self.wsc[4], self.C.g("fc.4.weight"))
        K.colsum(dlog, self.stat[4], self.C.g("fc.4.bias"))
        K.conv_dgrad(dlog, *self.geom[4][:4], self.wd[4], *self.geom[4][4:], self.df1d)
        K.binary(self.df1d, self.m1, K.MUL, self.df1d)                         # through Dropout
        K.unary_bwd(self.df1d, self.a[3].view(B, 256), K.RELU, self.d[3].view(B, 256))
        # fc.1 (7x7 convolution over the dropped-out map)
        d3 = self.d[3].view(B, 256)
        K.conv_wgrad(self.a3d, d3, *self.geom[3], self.wsc[3], self.C.g("fc.1.weight").view(256, 128, 7, 7))
        K.colsum(d3, self.stat[3], self.C.g("fc.1.bias"))
        K.conv_dgrad(d3, *self.geom[3][:4], self.wd[3], *self.geom[3][4:], self.da3d)
        K.binary(self.da3d, self.m2, K.MUL, self.da3d)                         # through Dropout2d
        K.unary_bwd(self.da3d, self.a[2], K.RELU, self.d[2])
        # conv stack
        for i in (2, 1, 0):
            xin = self.x if i == 0 else self.a[i - 1]
            K.conv_wgrad(xin, self.d[i], *self.geom[i], self.wsc[i], self.C.g(self.wname[i] + ".weight"))
            K.colsum(self.d[i].view(-1, self.geom[i][4]), self.stat[i], self.C.g(self.wname[i] + ".bias"))
            if i > 0:
                K.conv_dgrad(self.d[i], *self.geom[i][:4], self.wd[i], *self.geom[i][4:], self.d[i - 1], act_ref=self.a[i - 1],
                             ref_act=K.ACT_RELU)
        self.C.adam_step(self.lr)                                              # optim.Adam(lr=cls_lr), trainer.py:9
        self.refresh()

    # ------------------------------------------------------------------ API
    def step(self, x, y, masks=None):
        """One training iteration on a batch [B,1,28,28] / [B]; returns the loss (device scalar)."""
        self.x.view(-1).copy_(x.reshape(-1), non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if masks is not None:
            self.inject = True
            self.m2.copy_(masks[0].view(self.B, 1, 1, 128).expand(self.B, 7, 7, 128))
            self.m1.copy_(masks[1])
        if not self.use_graph:
            self._body()
            return self.loss
        if self.graph is None:
            snap = [t.clone() for t in (self.C.data, self.C.m, self.C.v, self.C.step, self.rng)]
            self._body()
            torch.cuda.synchronize()
            for dst, src in zip((self.C.data, self.C.m, self.C.v, self.C.step, self.rng), snap):
                dst.copy_(src)
            self.refresh()
            torch.cuda.synchronize()
            self.program = dataflow.record(self._body)
            self.graph = graphs.capture(self.program.emit)
        self.graph.replay()
        return self.loss

    def logits(self, x):
        """Eval-mode forward (dropout inactive)."""
        self.x.view(-1).copy_(x.reshape(-1), non_blocking=True)
        self._fwd(False)
        return self.a[4].view(self.B, 10)


def train_classifier(classifier, train_loader, valid_loader, cfg, device):
    """Drop-in for trainer.py:8-39."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pcg_b200.mnist.classifier_trainer needs a CUDA device (there is no CPU fallback)")
    classifier.to(device)
    plans = {}

    def plan_for(bs):
        p = plans.get(bs)
        if p is None:
            first = next(iter(plans.values()), None)
            p = plans[bs] = ClassifierPlan(bs, device, cfg.cls_lr)
            if first is None:
                p.adopt(classifier)
            else:                                   # one optimizer state for the whole run (tail batches)
                p.C = first.C
                p.refresh()
        return p

    best_acc, last = 0.0, None
    save_path = cfg.classifier_path
    for epoch in range(cfg.num_epochs_clf):
        for x, y in train_loader:
            p = plan_for(x.size(0))
            if last is not None and last is not p:
                p.refresh()                         # the other plan's Adam steps changed the weights
            p.step(x.to(device, non_blocking=True).float(), y.to(device, non_blocking=True).long())
            last = p
        correct, total = 0, 0
        for x, y in valid_loader:
            p = plan_for(x.size(0))
            p.refresh()
            preds = p.logits(x.to(device).float()).argmax(1)
            correct += (preds == y.to(device)).sum().item()
            total += y.size(0)
            last = p
        acc = correct / max(total, 1)
        print(f"[Classifier] Epoch {epoch+1}/{cfg.num_epochs_clf} | Val Acc: {acc:.4f}")
        if acc > best_acc:
            best_acc = acc
            os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
            torch.save(classifier.state_dict(), save_path)
    print(f"Saved best classifier with acc={best_acc:.4f} to {save_path}")
    return best_acc
