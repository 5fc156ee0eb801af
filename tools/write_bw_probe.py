import torch
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
for mb in (126, 252, 1024, 4096):
    x=torch.empty(mb*1024*1024//4, device='cuda'); y=torch.empty_like(x)
    tf=t(lambda: x.fill_(1.0)); tc=t(lambda: y.copy_(x))
    print(f"{mb} MB: fill {tf*1e3:.1f} us = {mb*1.048576/tf:.0f} GB/s write-only; copy {tc*1e3:.1f} us = {2*mb*1.048576/tc:.0f} GB/s r+w")
