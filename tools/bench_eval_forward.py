"""Inference forward of the MNIST generator (eval mode): samples/s of G.eval()(x, target, mask) at
batch 512 with the BatchNorms folded into the convolutions, next to the train-mode forward of the same module."""
import json, sys
sys.path.insert(0, '.')
import torch
import pcg_b200
from pcg_b200.mnist.models.generator import ResidualGenerator
from oracle import mnist_countergan as O
B = 512
torch.manual_seed(0)
G = ResidualGenerator().cuda()
x, y, t, m = (v.cuda().contiguous() for v in O.synth_batch(B, 5))
out = {}
with torch.no_grad():
    for mode in ("train", "eval"):
        G.train(mode == "train")
        for _ in range(5): G(x, t, m)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): G(x, t, m)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        out[mode] = {"ms_per_batch": ms, "samples_per_s": B / (ms * 1e-3)}
out["algorithmic_gflop_per_batch"] = 2 * 377.5e6 * B / 1e9
out["eval_tflops"] = out["algorithmic_gflop_per_batch"] / out["eval"]["ms_per_batch"]
print(json.dumps(out))
