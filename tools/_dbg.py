import sys, ctypes; sys.path.insert(0,'/root/repo')
import torch, numpy as np
import montage_gan_b200
from montage_gan_b200 import synth, _lib
lib=_lib.load(); dev=torch.device('cuda',0)
B,L,H,W=64,7,256,256; dt=1
P=lambda t: ctypes.c_void_p(t.data_ptr())
x=synth.make_layers(8,L,H,W,"S",seed=0).repeat(8,1,1,1,1).to(dev,torch.bfloat16).contiguous()
go=synth.make_grad_out(B,H,W,seed=0).to(dev,torch.bfloat16)
out=torch.empty(B,4,H,W,dtype=torch.bfloat16,device=dev); gx=torch.empty_like(x); gt=torch.empty(B,L,2,3,device=dev)
sav=torch.empty(lib.mgr_saved_alpha_bytes(B,L,H,W,dt),dtype=torch.uint8,device=dev)
wsb=lib.mgr_render_backward_workspace_bytes(B,L,H,W,dt,1,3); ws=torch.empty(wsb,dtype=torch.uint8,device=dev)
sp=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for seed in range(6):
    thc=synth.make_theta(B,L,"I",seed=seed); th=thc.to(dev)
    det=(thc[...,0,0]*thc[...,1,1]-thc[...,0,1]*thc[...,1,0]).abs().flatten()
    def f(): lib.mgr_render_forward(P(x),None,P(th),P(out),P(sav),B,L,H,W,dt,0,sp)
    def b(): lib.mgr_render_backward(P(x),None,P(th),P(out),P(go),P(sav),P(gx),P(gt),P(ws),wsb,B,L,H,W,dt,0,3,sp)
    f(); b(); torch.cuda.synchronize()
    res={}
    for name,fn in (('fwd',f),('bwd',b)):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize(); res[name]=e0.elapsed_time(e1)/5*1e3
    srt=np.sort(det.numpy())
    print(seed, {k:round(v) for k,v in res.items()}, 'smallest |det|', np.round(srt[:5],3), 'n<0.2:', int((srt<0.2).sum()))
