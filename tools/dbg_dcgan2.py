import sys; sys.path.insert(0, '.')
import torch
from oracle import dcgan as O
import pcg_b200
from pcg_b200.dcgan import DcganPlan
from pcg_b200 import ops as K
def rel(a,b):
    a,b=a.detach().float().cpu(),b.detach().float().cpu(); return ((a-b).abs().max()/(b.abs().max()+1e-30)).item()
B=8
PG, PD = O.synth_params(O.g_shapes(), 5), O.synth_params(O.d_shapes(), 6)
S = O.make_state(PG, O.buffers(O.g_shapes()), PD, O.buffers(O.d_shapes()))
plan = DcganPlan(B, "cuda", use_graph=False)
plan.G.load(PG); plan.D.load(PD); plan.refresh()
real, noise = O.synth_batch(B, 70)
D, DB = S["D"], S["DB"]
taps = {}
out = O.d_forward(D, DB, real, taps=taps)
for t in taps.values(): t.retain_grad()
err = O.bce(out, torch.ones(B))
g = dict(zip(D.keys(), torch.autograd.grad(err, list(D.values()), retain_graph=True)))
ga = torch.autograd.grad(err, [taps['d1'], taps['d2'], taps['d3']])
plan.real.view(-1).copy_(real.cuda().reshape(-1))
plan._d_fwd(plan.real, 0)
K.gan_loss(plan.dy[0][4].view(-1), K.GAN_BCE, 1.0, plan.scal[2:3], plan.dz, out_aux=plan.scal[4:5])
plan._d_bwd(plan.real, 0, lambda n: plan.D.g(n), True, False)
torch.cuda.synchronize()
print('loss', plan.scal[2].item(), err.item())
for i in (1,2,3):
    print('act d%d'%i, rel(plan.da[0][i].permute(0,3,1,2), taps['d%d'%i]))
for i,t in zip((1,2,3), ga):
    print('grad wrt a%d'%i, rel(plan.dd[i].permute(0,3,1,2), t))
for k in g: print('dD(real only)', k, rel(plan.D.g(k), g[k]), g[k].abs().max().item())
# isolate BN backward of D layer 1 with the plan's own tensors
import torch.nn.functional as F
y = plan.dy[0][1].reshape(-1, 128).clone().requires_grad_(True)
gam = plan.D.p("main.3.weight").clone().requires_grad_(True); bet = plan.D.p("main.3.bias").clone().requires_grad_(True)
o = F.leaky_relu(F.batch_norm(y, None, None, gam, bet, True, 0.1, 1e-5), 0.2)
o.backward(plan.dd[1].reshape(-1, 128))
print('iso BN1: a1', rel(o, plan.da[0][1].reshape(-1,128)), 'dy', rel(plan.ddy[1].reshape(-1,128), y.grad), 'dgamma', rel(plan.D.g("main.3.weight"), gam.grad), 'dbeta', rel(plan.D.g("main.3.bias"), bet.grad))
print('oracle dgamma vs torch-gpu dgamma', rel(g["main.3.weight"], gam.grad), rel(g["main.3.bias"], bet.grad))
st = plan.d_bn[1]["st"][0]
print('mean', rel(st.mean, y.detach().mean(0)), 'rstd', rel(st.rstd, 1/torch.sqrt(y.detach().var(0, unbiased=False)+1e-5)))
mine = plan.ddy[1].reshape(-1,128); ref = y.grad
d = (mine-ref).abs()
print('dy abs max', ref.abs().max().item(), 'err max', d.max().item(), 'frac rows with err>1e-3*max', (d.max(1).values > 1e-3*ref.abs().max()).float().mean().item())
bad_rows = torch.nonzero(d.max(1).values > 1e-3*ref.abs().max()).flatten()
print('bad rows (first 40):', bad_rows[:40].tolist(), 'count', bad_rows.numel())
bad_cols = torch.nonzero(d.max(0).values > 1e-3*ref.abs().max()).flatten()
print('bad cols:', bad_cols[:40].tolist(), 'count', bad_cols.numel())
# rerun the backward alone, twice, to see if it is deterministic / depends on stale scratch
dy2 = torch.empty_like(plan.ddy[1]); dg=torch.empty(128,device='cuda'); db=torch.empty(128,device='cuda')
K.bn_train_bwd(plan.dd[1], plan.dy[0][1], 2048, 128, plan.D.p("main.3.weight"), st, dy2, dg, db, act=K.ACT_LRELU, slope=0.2)
torch.cuda.synchronize()
print('rerun alone: dy', rel(dy2.reshape(-1,128), ref), 'dbeta', rel(db, bet.grad))
