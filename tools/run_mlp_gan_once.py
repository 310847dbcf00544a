import sys; sys.path.insert(0, '.')
import torch, bench
step, _, _ = bench._other_setup("cgan_moons", 1024, torch.device("cuda"))
for i in range(6): step(i)
torch.cuda.synchronize()
