import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
import montage_gan_b200
from montage_gan_b200 import render as mr, synth, _lib
from oracle import restatement as R
lib=_lib.load()
B,L,H,W=1,2,33,257
x=synth.make_layers(B,L,H,W,"W",seed=5); th=synth.make_theta(B,L,"I",seed=5)
ref=R.render_fwd(x.numpy(),th.numpy(),"m11",np.float64)
xc,tc=x.cuda(),th.cuda()
for rep in range(3):
  for path in (1,0):
    lib.mgr_set_debug_path(path); a=mr.render(xc,tc).cpu().numpy()
    d=np.abs(a-ref).max(axis=(0,1)); bad=np.argwhere(d>1e-4)
    print('rep',rep,'path',path,'nbad',len(bad), bad[:6].tolist(), 'maxerr',d.max())
lib.mgr_set_debug_path(0)
xx=xc.clone().requires_grad_(True); tt=tc.clone().requires_grad_(True)
o=mr.render(xx,tt); o.backward(torch.randn_like(o)); torch.cuda.synchronize()
a=o.detach().cpu().numpy(); d=np.abs(a-ref).max(axis=(0,1)); bad=np.argwhere(d>1e-4)
print('with grad: nbad',len(bad), bad[:6].tolist(), 'maxerr',d.max())
