import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200
from montage_gan_b200 import synth
from montage_gan_b200.host import HostRenderer
B,L,H,W=64,7,256,256
dt=torch.bfloat16
x=synth.make_layers(8,L,H,W,"S",seed=0).repeat(8,1,1,1,1).to(dt).pin_memory()
th=synth.make_theta(B,L,"I",seed=0).pin_memory(); go=synth.make_grad_out(B,H,W,seed=0).to(dt).pin_memory()
for chunk in (2,4,8,16,32,64):
    hr=HostRenderer(B,L,H,W,dt,chunk_B=chunk)
    for _ in range(2): hr.fwd_bwd(x,th,go)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): hr.fwd_bwd(x,th,go)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(json.dumps({"chunk":chunk,"ms":round(ms,3),"Mpix_s":round(B*L*H*W/1e3/ms),"GBs_each_way":round(268.4/ms,1)}))
    del hr
