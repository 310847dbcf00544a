"""Throughput of the on-device MNIST input pipeline (one shuffled epoch of 54,000 synthetic uint8 images, batch 512)
next to the reference's DataLoader path (ToTensor + Normalize on PIL images in worker processes) on the host cores."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
import pcg_b200
from pcg_b200.mnist import data_utils as DU
N, B = 54000, 512
g = torch.Generator().manual_seed(0)
u8 = torch.randint(0, 256, (N, 28, 28), generator=g, dtype=torch.uint8)
y = torch.randint(0, 10, (N,), generator=g)
loader = DU.DeviceLoader(u8.cuda(), y.cuda(), None, B, shuffle=True)
for _ in loader: pass
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
nb = 0
for x, yy in loader: nb += 1
e1.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
out = {"metric": "MNIST input pipeline samples/s (shuffled epoch, batch 512)", "native_device_ms": e0.elapsed_time(e1),
       "native_samples_per_s_device": N / (e0.elapsed_time(e1) * 1e-3), "native_samples_per_s_wall": N / wall, "batches": nb,
       "bytes_per_sample": 784 + 784 * 4 + 16}
try:
    from torchvision import transforms
    from PIL import Image
    t = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,), (0.5,))])
    class DS(torch.utils.data.Dataset):
        def __len__(self): return 8192
        def __getitem__(self, i): return t(Image.fromarray(u8[i].numpy(), mode="L")), int(y[i])
    dl = torch.utils.data.DataLoader(DS(), batch_size=128, shuffle=True, num_workers=4)
    for _ in dl: break
    t0 = time.perf_counter(); n = 0
    for xb, yb in dl: n += xb.shape[0]
    out["reference_dataloader_samples_per_s"] = n / (time.perf_counter() - t0)
    out["reference_sample"] = "torchvision transforms on PIL images, DataLoader(batch 128, 4 workers), 8192 images"
except Exception as e:
    out["reference_dataloader_samples_per_s"] = None; out["reference_error"] = str(e)
print(json.dumps(out))
