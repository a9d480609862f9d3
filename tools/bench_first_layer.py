import sys, torch
sys.path.insert(0,'/root/repo')
import camvid_b200
from camvid_b200 import ops
dev=torch.device('cuda')
x=torch.randn(16,3,360,480,device=dev)
cols=torch.empty(16,360,480,64,dtype=torch.bfloat16,device=dev)
w=torch.randn(64,3,3,3,device=dev)*0.2
wp=ops.pack_weights_fprop(w,1,64,64)
y=torch.empty(16,360,480,64,dtype=torch.bfloat16,device=dev)
parts=torch.empty(ops.stat_rows(),2,64,device=dev)
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
for name,fn in (("im2col",lambda: ops.im2col3x3(x,cols)),("conv taps1",lambda: ops.conv3x3(cols,wp,y,taps=1,stat_partials=parts))):
    for _ in range(2): fn()
    ts=[]
    for _ in range(7):
        flush.zero_(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(name, sorted(ts)[3]*1e3, "us")
ref=torch.nn.functional.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), padding=1)
got=y.float().permute(0,3,1,2)
print("rel", ((got-ref).norm()/ref.norm()).item(), "stat rel", ((parts.sum(0)[0]-ref.sum((0,2,3))).norm()/ref.sum((0,2,3)).norm()).item())
