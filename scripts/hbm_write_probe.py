import torch
x=torch.empty(4300*1024*1024//2,dtype=torch.bfloat16,device='cuda')
y=torch.empty_like(x)
for fn,name,b in ((lambda: x.zero_(),'memset 4.3GB',x.numel()*2),(lambda: y.copy_(x),'copy 4.3GB (r+w)',x.numel()*4)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(name, f"{ms:.3f} ms  {b/ms/1e6:.0f} GB/s")
