import numpy as np, torch, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import montage_gan_b200
from montage_gan_b200 import synth
from oracle import restatement as R
from test_tma_stencil_gpu import _theta, _run
for (B,L,H,W) in [(3,7,80,136),(1,7,80,136),(1,7,64,128),(1,7,60,128)]:
    x = synth.make_layers(B, L, H, W, "S", seed=72).to(torch.bfloat16).float()
    th=_theta(B,L,72,0.7)
    go = synth.make_grad_out(B, H, W, seed=72).to(torch.bfloat16).float()
    new=_run(x,th,go,"m11",torch.bfloat16,0); old=_run(x,th,go,"m11",torch.bfloat16,4)
    r64 = R.render_fwd_bwd(x.numpy(), th.numpy(), go.numpy(), "m11", np.float64)
    en=np.abs(new[2]-r64["grad_theta"]).reshape(B,L,6); eo=np.abs(old[2]-r64["grad_theta"]).reshape(B,L,6)
    print((B,L,H,W),"max err new",en.max(),"old",eo.max(),"max ref",np.abs(r64["grad_theta"]).max())
    idx=np.unravel_index(en.argmax(),en.shape); print(" worst",idx,"new",new[2].reshape(B,L,6)[idx],"ref",r64["grad_theta"].reshape(B,L,6)[idx],"old",old[2].reshape(B,L,6)[idx])
    print(" theta px shift of worst layer", th[idx[0],idx[1],:,2].numpy()*np.array([W/2,H/2]))
    print(" err by layer", en.max(axis=2).round(2))
