import torch, time
torch.cuda.init()
n = 6*1024**3
x = torch.empty(n, dtype=torch.uint8, device='cuda')
y = torch.empty(n, dtype=torch.uint8, device='cuda')
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
ms=t(lambda: x.zero_()); print('memset write-only GB/s', n/ms/1e6)
xf = x.view(torch.float32)
ms=t(lambda: xf.fill_(1.5)); print('fill f32 write-only GB/s', n/ms/1e6)
ms=t(lambda: y.copy_(x)); print('copy GB/s (r+w)', 2*n/ms/1e6)
ms=t(lambda: xf.sum()); print('read-only sum GB/s', n/ms/1e6)
