import sys; sys.path.insert(0, '.')
import torch, torch.nn.functional as F
import pcg_b200
from pcg_b200 import ops as K
torch.manual_seed(0)
dev='cuda'
def rel(a,b): return ((a-b).abs().max()/(b.abs().max()+1e-30)).item()
for (M,C) in [(2048,128),(512,256),(128,512),(8192,64)]:
    y = torch.randn(M,C,device=dev)*1.3+0.2; gamma=torch.rand(C,device=dev)+0.5; beta=torch.randn(C,device=dev)*0.1
    dz = torch.randn(M,C,device=dev)
    yr = y.clone().requires_grad_(True); g_=gamma.clone().requires_grad_(True); b_=beta.clone().requires_grad_(True)
    out = F.leaky_relu(F.batch_norm(yr, None, None, g_, b_, True, 0.1, 1e-5), 0.2)
    out.backward(dz)
    st = K.BNState(C, dev); z = torch.empty_like(y); rm=torch.zeros(C,device=dev); rv=torch.ones(C,device=dev); nbt=torch.zeros((),dtype=torch.int64,device=dev)
    K.bn_train_fwd(y, M, C, gamma, beta, rm, rv, nbt, st, z, act=K.ACT_LRELU, slope=0.2)
    dy=torch.empty_like(y); dg=torch.empty(C,device=dev); db=torch.empty(C,device=dev)
    K.bn_train_bwd(dz, y, M, C, gamma, st, dy, dg, db, act=K.ACT_LRELU, slope=0.2)
    torch.cuda.synchronize()
    print('bn', M, C, 'fwd', rel(z,out), 'dy', rel(dy, yr.grad), 'dgamma', rel(dg,g_.grad), 'dbeta', rel(db,b_.grad))
# conv k4 s2 p1 dgrad / wgrad / fprop
for (B,H,Ci,Co) in [(8,32,64,128),(8,16,128,256),(8,8,256,512),(8,64,1,64)]:
    x = torch.randn(B,Ci,H,H,device=dev); w = torch.randn(Co,Ci,4,4,device=dev)*0.05
    xr = x.clone().requires_grad_(True); wr=w.clone().requires_grad_(True)
    yo = F.conv2d(xr, wr, None, 2, 1); dyo = torch.randn_like(yo); yo.backward(dyo)
    xn = x.permute(0,2,3,1).contiguous(); dyn = dyo.permute(0,2,3,1).contiguous()
    wf = torch.empty(Co*Ci*16,device=dev); wd=torch.empty(Co*Ci*16,device=dev); K.pack_weights(w,4,wf=wf,wd=wd)
    Ho=H//2
    out = torch.empty(B,Ho,Ho,Co,device=dev); K.conv_fprop(xn,B,H,H,Ci,wf,Co,4,2,1,out)
    din = torch.empty(B,H,H,Ci,device=dev); K.conv_dgrad(dyn,B,H,H,Ci,wd,Co,4,2,1,din)
    sc = K.conv_wgrad_scratch(B,H,H,Ci,Co,4,2,1,dev); dw=torch.empty(Co,Ci,4,4,device=dev); K.conv_wgrad(xn,dyn,B,H,H,Ci,Co,4,2,1,sc,dw)
    torch.cuda.synchronize()
    print('conv', B,H,Ci,Co,'fprop', rel(out.permute(0,3,1,2), yo), 'dgrad', rel(din.permute(0,3,1,2), xr.grad), 'wgrad', rel(dw, wr.grad))
