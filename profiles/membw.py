import torch, time
dev='cuda:0'
n=1<<30  # 4 GiB fp32
x=torch.empty(n,dtype=torch.float32,device=dev); y=torch.empty(n,dtype=torch.float32,device=dev)
def t(f,reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); b=torch.cuda.Event(True); a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
ms=t(lambda: x.fill_(1.0)); print('write  GB/s', 4*n/ms/1e6)
ms=t(lambda: x.sum()); print('read   GB/s', 4*n/ms/1e6)
ms=t(lambda: y.copy_(x)); print('copy   GB/s (r+w)', 8*n/ms/1e6)
h=x.view(torch.int32)
# 1 read : 3 write pattern like QKV gemm: out[3n/4]... emulate with cat of 3 copies into bigger buffer
z=torch.empty(3*(n//4),dtype=torch.float32,device=dev); xs=x[:n//4]
ms=t(lambda: (z[:n//4].copy_(xs), z[n//4:2*(n//4)].copy_(xs), z[2*(n//4):].copy_(xs))); print('1r(L2 reuse?):3w GB/s', (4*(n//4)*4)/ms/1e6,'(counting 1 read + 3 writes)')
