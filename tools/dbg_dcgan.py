import sys; sys.path.insert(0, '.')
import torch
from oracle import dcgan as O
import pcg_b200
from pcg_b200.dcgan import DcganPlan
def rel(a,b):
    a,b=a.detach().float().cpu(),b.detach().float().cpu(); return ((a-b).abs().max()/(b.abs().max()+1e-30)).item()
B=int(sys.argv[1]) if len(sys.argv)>1 else 8
PG, PD = O.synth_params(O.g_shapes(), 5), O.synth_params(O.d_shapes(), 6)
S = O.make_state(PG, O.buffers(O.g_shapes()), PD, O.buffers(O.d_shapes()))
plan = DcganPlan(B, "cuda", use_graph=False)
plan.G.load(PG); plan.D.load(PD); plan.refresh()
real, noise = O.synth_batch(B, 70)
# oracle with taps
import torch.nn.functional as F
G, GB, D, DB = S["G"], S["GB"], S["D"], S["DB"]
sc, gr = O.dcgan_step(S, real, noise)
got = plan.step(real.cuda(), noise.cuda()).tolist()
print('scalars', got[:7], sc)
for k in gr["G"]: print('dG', k, rel(plan.G.g(k), gr["G"][k]))
for k in gr["D"]: print('dD', k, rel(plan.D.g(k), gr["D"][k]))
for k in S["D"]:
    d = (plan.D.p(k).cpu() - S["D"][k].detach()).abs()
    print('pD', k, 'max', d.max().item()/2e-4, 'mean', d.mean().item()/2e-4, 'frac>0.5lr', (d > 1e-4).float().mean().item())
