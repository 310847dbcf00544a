"""Micro-benchmark + error report of the one-channel 4x4/stride-2 kernels at the DCGAN shape (N=256, 64x64)."""
import sys
sys.path.insert(0, '.')
import torch
import torch.nn.functional as F
import pcg_b200  # noqa: F401
from pcg_b200 import ops as K

N, HW = 256, 64
Ho = HW // 2
torch.manual_seed(0)
w = torch.randn(64, 1, 4, 4, device="cuda") * 0.2
x = torch.randn(N, 1, HW, HW, device="cuda")
wf, wd = torch.empty(1024, device="cuda"), torch.empty(1024, device="cuda")
K.pack_weights(w, 4, wf=wf, wd=wd)
xn = x.permute(0, 2, 3, 1).contiguous()
dy = torch.randn(N, 64, Ho, Ho, device="cuda")
dyn = dy.permute(0, 2, 3, 1).contiguous()
out = torch.empty(N, Ho, Ho, 64, device="cuda")
dx = torch.empty(N, HW, HW, 1, device="cuda")
dw = torch.empty(64, 1, 4, 4, device="cuda")
scratch = K.conv_wgrad_scratch(N, HW, HW, 1, 64, 4, 2, 1, "cuda")
ops = {"fprop": lambda: K.conv_fprop(xn, N, HW, HW, 1, wf, 64, 4, 2, 1, out, act=K.ACT_LRELU),
       "dgrad": lambda: K.conv_dgrad(dyn, N, HW, HW, 1, wd, 64, 4, 2, 1, dx),
       "wgrad": lambda: K.conv_wgrad(xn, dyn, N, HW, HW, 1, 64, 4, 2, 1, scratch, dw)}
from pcg_b200 import graphs
for name, fn in ops.items():
    for _ in range(3):
        fn()
    g = graphs.capture(lambda: [fn() for _ in range(20)])      # replayed graph: no host launch overhead in the timing
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (graph of 20)")
rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
print("fprop err", rel(out.permute(0, 3, 1, 2), F.leaky_relu(F.conv2d(x, w, None, 2, 1), 0.2)))
ref = F.conv_transpose2d(dy, w, None, 2, 1)
print("dgrad err", rel(dx.view(N, 1, HW, HW), ref))
d = (dx.view(N, 1, HW, HW) - ref).abs()
print("dgrad worst at", [int(v) for v in torch.unravel_index(d.argmax(), d.shape)], "border rows err", d[:, :, 0].max().item(), d[:, :, -1].max().item(),
      "border cols", d[:, :, :, 0].max().item(), d[:, :, :, -1].max().item(), "interior", d[:, :, 8:56, 8:56].max().item())
wz = w.clone().requires_grad_(True)
(gw,) = torch.autograd.grad(F.conv2d(x, wz, None, 2, 1), wz, dy)
print("wgrad err", rel(dw, gw))
