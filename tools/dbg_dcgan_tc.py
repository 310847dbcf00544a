import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from oracle import dcgan as O
import pcg_b200
from pcg_b200.dcgan import DcganPlan
def l2(a,b):
    a,b=a.detach().double().cpu(),b.detach().double().cpu(); return ((a-b).norm()/(b.norm()+1e-300)).item()
B=int(sys.argv[1]) if len(sys.argv)>1 else 16
PG, PD = O.synth_params(O.g_shapes(), 5), O.synth_params(O.d_shapes(), 6)
S = O.make_state(PG, O.buffers(O.g_shapes()), PD, O.buffers(O.d_shapes()))
plan = DcganPlan(B, "cuda", use_graph=False, tensor_cores=True)
plan.G.load(PG); plan.D.load(PD); plan.refresh()
real, noise = O.synth_batch(B, 170)
sc, gr = O.dcgan_step(S, real, noise)
plan.real.view(-1).copy_(real.cuda().reshape(-1)); plan.noise.view(-1).copy_(noise.cuda().reshape(-1))
from pcg_b200 import ops as K
K.set_conv_tensor_cores(True)
plan._d_phase(); torch.cuda.synchronize()
print("D grads:", sorted(((round(l2(plan.D.g(k), gr["D"][k]),4), k) for k in gr["D"]), reverse=True)[:6])
plan.D.load({k: v.detach() for k, v in S["D"].items()}); plan.refresh()
plan._g_phase(); torch.cuda.synchronize()
K.set_conv_tensor_cores(False)
print("G grads:", sorted(((round(l2(plan.G.g(k), gr["G"][k]),4), k) for k in gr["G"]), reverse=True)[:6])
print("dfake:", l2(plan.dfake.view(-1), gr["dfake"].reshape(-1)) if "dfake" in gr else "n/a")
