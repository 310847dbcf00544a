"""How far do the DCGAN loss curves of the native plan drift from the fp32 CPU oracle over N iterations, per operand
precision of the 64..512-channel convolutions?   python tools/exp_dcgan_precision.py [steps] [batch]
Modes: fp32 (CUDA cores), bf16x3 (tensor cores, fp32-equivalent products), bf16 (tensor cores, plain bf16 operands).
The fp32 column is the spread between two fp32 realisations of the same iteration (different summation orders): the
yardstick for the other two."""
import sys
sys.path.insert(0, '.')
import torch
from oracle import dcgan as O
import pcg_b200  # noqa: F401
from pcg_b200 import ops as K
from pcg_b200.dcgan import DcganPlan

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
PG, PD = O.synth_params(O.g_shapes(), 5), O.synth_params(O.d_shapes(), 6)
S = O.make_state(PG, O.buffers(O.g_shapes()), PD, O.buffers(O.d_shapes()))
batches = [O.synth_batch(B, 1000 + i) for i in range(steps)]
keys = ("errD", "errG", "D_x", "D_G_z1")
ora = []
for b in batches:
    sc, _ = O.dcgan_step(S, *b)
    ora.append([sc[k] for k in keys])
ora = torch.tensor(ora).double()
pG = torch.cat([v.detach().flatten() for v in S["G"].values()]).double()
pG0 = torch.cat([v.flatten() for v in PG.values()]).double()
for name, tc, terms in (("fp32", False, 3), ("bf16x3", True, 3), ("bf16", True, 1)):
    K.set_conv_tensor_core_terms(terms)
    plan = DcganPlan(B, "cuda", use_graph=False, tensor_cores=tc)
    plan.G.load(PG); plan.D.load(PD); plan.refresh()
    nat = []
    for b in batches:
        nat.append(plan.step(b[0].cuda(), b[1].cuda()).clone())
    torch.cuda.synchronize()
    nat = torch.stack(nat).cpu().double()[:, [0, 1, 4, 5]]
    rel = (nat - ora).abs() / ora.abs().clamp_min(1e-3)
    g = torch.cat([plan.G.p(k).detach().flatten().cpu() for k in PG]).double()
    a, b_ = g - pG0, pG - pG0
    cos = (a @ b_ / (a.norm() * b_.norm())).item()
    for upto in (10, 50, 100, steps):
        print(f"{name:7s} steps<= {upto:4d}  max rel dev " + "  ".join(f"{k}={rel[:upto, i].max():.3e}" for i, k in enumerate(keys)))
    print(f"{name:7s} generator movement after {steps} steps: cos={cos:.4f} |native|/|oracle|={(a.norm() / b_.norm()).item():.4f}")
K.set_conv_tensor_core_terms(3)
