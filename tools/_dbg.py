import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
import montage_gan_b200
from montage_gan_b200 import render as mr, synth, _lib
from oracle import restatement as R
th=synth.make_theta(1,32,"I",seed=5)[:,23:26].contiguous()
H=W=16
x=synth.make_layers(1,3,H,W,"W",seed=5)
xn=x.numpy(); thn=th.numpy()
t=thn[0,1]
bad=[]
for i in range(8,16):
  for j in range(8,16):
    g2=torch.zeros(1,4,H,W); g2[0,:,i,j]=torch.tensor([1.,-2.,0.5,1.5])
    r=R.render_fwd_bwd(xn,thn,g2.numpy(),"m11",np.float64)
    xx=x.cuda().requires_grad_(True); tt=th.cuda().requires_grad_(True)
    o=mr.render(xx,tt); o.backward(g2.cuda()); torch.cuda.synchronize()
    gt=tt.grad.cpu().numpy().reshape(3,6)[1]; rf=r['grad_theta'].reshape(3,6)[1]
    e=np.abs(gt-rf).max()/max(1e-9,np.abs(rf).max())
    if e>1e-4:
        ix=t[0,0]*(j+.5-8)+t[0,1]*(i+.5-8)+t[0,2]*8+7.5; iy=t[1,0]*(j+.5-8)+t[1,1]*(i+.5-8)+t[1,2]*8+7.5
        bad.append((i,j,round(float(ix),3),round(float(iy),3),gt[3:].round(4).tolist(),rf[3:].round(4).tolist()))
print(len(bad)); 
for b in bad[:12]: print(b)
