"""Native mirror of the KC house-sales classifier pre-training, ``house_sales_kc_usa/trainer.py:18-180`` (SURVEY.md 8f
row 3).  ``train_classifier(X_train_all, X_test, y_train_all, y_test, scaler, config)`` keeps the reference's signature,
its stratified validation split, balanced class weights, AdamW + ReduceLROnPlateau(factor 0.5, patience 4) + early
stopping logic, checkpoint files and prints; the training iteration and the validation pass run in libpcg:
Linear -> LeakyReLU(0.1) -> train-mode BatchNorm1d -> Dropout (nn_classifier.py:7-28), class-weighted cross entropy,
backward, AdamW - primitive operators on fp32 tensors, replayed as a data-flow captured CUDA graph.  The learning rate
lives in device memory, so the scheduler changes it without re-capturing.
"""
import os

import numpy as np
import torch

from .. import dataflow
from .. import graphs
from .. import ops as K
from .kc import NNClassifier

DIMS = [256, 256, 128, 64]
P_DROP = [0.3, 0.2, 0.1, 0.0]
LIN = ["net.0", "net.4", "net.8", "net.12", "net.15"]
BN = ["net.2", "net.6", "net.10", "net.14"]


class KcClassifierPlan:
    def __init__(self, batch, device, input_dim=17, out_dim=4, lr=1e-3, wd=1e-4, class_weights=None, share=None):
        self.B, self.wd, self.d, self.nc = batch, wd, input_dim, out_dim
        dev = self.dev = torch.device(device)
        B = batch
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        dims = [input_dim] + DIMS
        names = []
        for j in range(4):
            names += [(LIN[j] + ".weight", (dims[j + 1], dims[j])), (LIN[j] + ".bias", (dims[j + 1],)),
                      (BN[j] + ".weight", (dims[j + 1],)), (BN[j] + ".bias", (dims[j + 1],))]
        names += [(LIN[4] + ".weight", (out_dim, 64)), (LIN[4] + ".bias", (out_dim,))]
        if share is not None:                       # tail batch of an epoch: same parameters, optimizer state, buffers
            self.C, self.rm, self.rv, self.nbt, self.lr_dev, self.cw, self.rng = (share.C, share.rm, share.rv, share.nbt,
                                                                                  share.lr_dev, share.cw, share.rng)
        else:
            self.C = K.FlatParams(names, dev)
            self.rm = [z(c) for c in DIMS]
            self.rv = [torch.ones(c, device=dev) for c in DIMS]
            self.nbt = [torch.zeros((), dtype=torch.int64, device=dev) for _ in DIMS]
            self.lr_dev = torch.full((1,), float(lr), device=dev)
            self.cw = None if class_weights is None else class_weights.to(dev).float().contiguous()
            self.rng = torch.zeros(3, dtype=torch.int64, device=dev)
            self.rng[2] = torch.initial_seed() & (2 ** 62 - 1)
        io = dims + [out_dim]
        self.wT = [z(io[j], io[j + 1]) for j in range(5)]
        self.wsc = [torch.zeros(K.conv_wgrad_scratch_floats(B, 1, 1, io[j], io[j + 1], 1, 1, 0) + 1024, device=dev) for j in range(5)]
        self.stat = [K.stat_scratch(max(io[j + 1], 4), dev) for j in range(5)]
        self.st = [K.BNState(c, dev) for c in DIMS]
        self.x = z(B, input_dim)
        self.y = torch.zeros(B, dtype=torch.int64, device=dev)
        self.a = [z(B, c) for c in DIMS]            # LeakyReLU(Linear)
        self.n = [z(B, c) for c in DIMS]            # BatchNorm output
        self.h = [z(B, c) for c in DIMS]            # after Dropout
        self.m = [z(B, c) for c in DIMS[:3]]        # keep-masks
        self.dh = [z(B, c) for c in DIMS]
        self.da = [z(B, c) for c in DIMS]
        self.du = [z(B, c) for c in DIMS]
        self.logits, self.dlogits = z(B, out_dim), z(B, out_dim)
        self.scal = z(4)                            # 0 loss, 1 correct
        self.inject = False
        self.graph, self.graph_eval = None, None
        self.refresh()

    def adopt(self, module):
        self.C.adopt(module)
        bns = [m for m in module.modules() if isinstance(m, torch.nn.BatchNorm1d)]
        for j, m in enumerate(bns):
            self.rm[j].copy_(m.running_mean)
            self.rv[j].copy_(m.running_var)
            self.nbt[j].copy_(m.num_batches_tracked)
            m._buffers["running_mean"], m._buffers["running_var"], m._buffers["num_batches_tracked"] = (
                self.rm[j], self.rv[j], self.nbt[j])
        self.refresh()

    def refresh(self):
        K.transpose_multi([(self.C.p(LIN[j] + ".weight"), self.wT[j]) for j in range(5)])

    def _fwd(self, train):
        h = self.x
        for j in range(4):
            K.linear_fwd(h, self.C.p(LIN[j] + ".weight"), self.a[j], self.C.p(LIN[j] + ".bias"), K.ACT_LRELU, 0.1)
            g, b = self.C.p(BN[j] + ".weight"), self.C.p(BN[j] + ".bias")
            if train:
                K.bn_train_fwd(self.a[j], self.B, DIMS[j], g, b, self.rm[j], self.rv[j], self.nbt[j], self.st[j], self.n[j])
            else:
                K.bn_eval(self.a[j], g, b, self.rm[j], self.rv[j], self.n[j])
            h = self.n[j]
            if train and j < 3:
                K.binary(self.n[j], self.m[j], K.MUL, self.h[j])
                h = self.h[j]
        K.linear_fwd(h, self.C.p(LIN[4] + ".weight"), self.logits, self.C.p(LIN[4] + ".bias"))
        return h

    def _body(self):
        if not self.inject:
            for j in range(3):
                K.dropout_mask(self.m[j], P_DROP[j], rng_state=self.rng)
        hlast = self._fwd(True)
        K.ce_loss_weighted(self.logits, self.y, self.cw, self.scal[0:1], self.dlogits, self.scal[1:2])
        K.linear_wgrad(hlast, self.dlogits, self.wsc[4], self.C.g(LIN[4] + ".weight"), self.C.g(LIN[4] + ".bias"), self.stat[4])
        K.linear_dgrad(self.dlogits, self.wT[4], self.dh[3], DIMS[3])
        d = self.dh[3]
        for j in (3, 2, 1, 0):
            if j < 3:
                K.binary(d, self.m[j], K.MUL, d)                               # through Dropout
            K.bn_train_bwd(d, self.a[j], self.B, DIMS[j], self.C.p(BN[j] + ".weight"), self.st[j], self.da[j],
                           self.C.g(BN[j] + ".weight"), self.C.g(BN[j] + ".bias"))
            K.unary_bwd(self.da[j], self.a[j], K.LRELU, self.du[j], 0.1)       # LeakyReLU(0.1): sign(a) = sign(pre-activation)
            xin = self.x if j == 0 else (self.h[j - 1] if j - 1 < 3 else self.n[j - 1])
            K.linear_wgrad(xin, self.du[j], self.wsc[j], self.C.g(LIN[j] + ".weight"), self.C.g(LIN[j] + ".bias"), self.stat[j])
            if j > 0:
                K.linear_dgrad(self.du[j], self.wT[j], self.dh[j - 1], DIMS[j - 1])
                d = self.dh[j - 1]
        K.adamw(self.C.data, self.C.grad, self.C.m, self.C.v, self.C.step, self.lr_dev, weight_decay=self.wd)
        self.refresh()

    def _eval_body(self):
        self._fwd(False)
        K.ce_loss_weighted(self.logits, self.y, self.cw, self.scal[2:3], None, self.scal[3:4])

    def _state(self):
        return [self.C.data, self.C.m, self.C.v, self.C.step, self.rng] + self.rm + self.rv + self.nbt

    def step(self, x, y, masks=None):
        """One training iteration; returns the scalar block (0 = loss, 1 = correct predictions)."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if masks is not None:
            self.inject = True
            for dst, src in zip(self.m, masks):
                dst.copy_(src)
        if self.graph is None:
            snap = [t.clone() for t in self._state()]
            self._body()
            torch.cuda.synchronize()
            for dst, src in zip(self._state(), snap):
                dst.copy_(src)
            self.refresh()
            torch.cuda.synchronize()
            self.graph = graphs.capture(dataflow.record(self._body).emit)
        self.graph.replay()
        return self.scal

    def evaluate(self, x, y):
        """Validation pass on one batch (eval-mode BatchNorm, no dropout): scal[2] = loss, scal[3] = correct."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.graph_eval is None:
            self._eval_body()
            torch.cuda.synchronize()
            self.graph_eval = graphs.capture(dataflow.record(self._eval_body).emit)
        self.graph_eval.replay()
        return self.scal


def train_classifier(X_train_all, X_test, y_train_all, y_test, scaler, config):
    """Drop-in for house_sales_kc_usa/trainer.py:18-180 (plots omitted when matplotlib is absent)."""
    from sklearn.model_selection import train_test_split
    from sklearn.utils.class_weight import compute_class_weight
    device = torch.device(config['cuda'])
    if device.type != "cuda":
        raise RuntimeError("pcg_b200.tabular.kc_classifier needs a CUDA device (there is no CPU fallback)")
    seed = config.get('seed', 42)
    torch.manual_seed(seed)
    np.random.seed(seed)
    out_dir = os.path.join(config.get('out_dir', '.'), "classifier_eval")
    os.makedirs(out_dir, exist_ok=True)
    X_train, X_val, y_train, y_val = train_test_split(X_train_all, y_train_all, test_size=config.get('val_frac', 0.10),
                                                      random_state=seed, stratify=y_train_all)
    num_classes = int(np.unique(y_train_all).size)
    config['num_classes'] = num_classes
    bs = config.get('clf_batch_size', config.get('batch_size', 128))
    Xt, yt = torch.tensor(X_train, dtype=torch.float32, device=device), torch.tensor(y_train, dtype=torch.long, device=device)
    Xv, yv = torch.tensor(X_val, dtype=torch.float32, device=device), torch.tensor(y_val, dtype=torch.long, device=device)
    model = NNClassifier(config['input_dim'], output_dim=num_classes).to(device)
    cw = torch.tensor(compute_class_weight('balanced', classes=np.arange(num_classes), y=y_train), dtype=torch.float32)
    lr = config.get('clf_lr', 1e-3)
    plans = {}

    def plan_for(n):
        p = plans.get(n)
        if p is None:
            first = next(iter(plans.values()), None)
            p = plans[n] = KcClassifierPlan(n, device, config['input_dim'], num_classes, lr, config.get('clf_wd', 1e-4), cw, share=first)
            if first is None:
                p.adopt(model)
        return p

    # ReduceLROnPlateau(mode='min', factor=0.5, patience=4) and early stopping, trainer.py:61-67,125-147
    best_val, best_state, wait, bad, cur_lr = float('inf'), None, 0, 0, lr
    sched_best = float('inf')
    patience, epochs = config.get('clf_early_stopping', 15), config.get('clf_epochs', 100)
    path = config.get('clf_model_path', os.path.join(out_dir, 'clf_model_best.pth'))
    last = None
    for epoch in range(1, epochs + 1):
        perm = torch.randperm(Xt.shape[0], device=device)
        acc = torch.zeros(2, device=device)
        total = 0
        for i in range(0, Xt.shape[0], bs):
            idx = perm[i:i + bs]
            p = plan_for(idx.numel())
            if last is not None and last is not p:
                p.refresh()
            sc = p.step(Xt[idx], yt[idx])
            acc += torch.stack([sc[0] * idx.numel(), sc[1]])
            total += idx.numel()
            last = p
        tl, tc = acc.tolist()
        train_loss, train_acc = tl / total, tc / total
        acc.zero_()
        vtotal = 0
        for i in range(0, Xv.shape[0], bs):
            n = min(bs, Xv.shape[0] - i)
            p = plan_for(n)
            if last is not p:
                p.refresh()
            sc = p.evaluate(Xv[i:i + n], yv[i:i + n])
            acc += torch.stack([sc[2] * n, sc[3]])
            vtotal += n
            last = p
        vl, vc = acc.tolist()
        val_loss, val_acc = vl / vtotal, vc / vtotal
        # scheduler.step(val_loss): torch's rel threshold 1e-4, patience 4, factor 0.5
        if val_loss < sched_best * (1 - 1e-4):
            sched_best, bad = val_loss, 0
        else:
            bad += 1
        if bad > 4:
            cur_lr *= 0.5
            next(iter(plans.values())).lr_dev.fill_(cur_lr)
            bad = 0
        if val_loss < best_val - 1e-6:
            best_val, wait = val_loss, 0
            best_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
            torch.save({'model_state_dict': best_state, 'epoch': epoch, 'scaler': scaler}, path)
        else:
            wait += 1
        print(f"[Epoch {epoch}/{epochs}] train_loss={train_loss:.4f}, train_acc={train_acc:.4f} | val_loss={val_loss:.4f}, "
              f"val_acc={val_acc:.4f} | wait={wait}")
        if wait >= patience:
            print(f"Early stopping at epoch {epoch} (no improvement for {patience} epochs).")
            break
    if best_state is not None:
        model.load_state_dict(best_state)
    final = config.get('clf_model_path', os.path.join(out_dir, 'clf_model_final.pth'))
    if os.path.dirname(final):
        os.makedirs(os.path.dirname(final), exist_ok=True)
    torch.save(model.state_dict(), final)
    print(f"Saved classifier model to {final}")
    print("Training finished. Best val loss: %.4f" % best_val)
    return model
