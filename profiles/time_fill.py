import torch
x = torch.empty(8,19,512,1024, device='cuda')
for _ in range(3): x.fill_(1.0)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.fill_(1.0)
e1.record(); torch.cuda.synchronize()
t=e0.elapsed_time(e1)/10
print(f"torch fill 319 MB: {t*1e3:.1f} us {x.numel()*4/t/1e6:.0f} GB/s")
