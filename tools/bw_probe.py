import torch, time
x = torch.empty(1<<30, dtype=torch.float32, device="cuda")  # 4 GiB
y = torch.empty(1<<30, dtype=torch.float32, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(n):
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_()); print("memset 4GiB: %.3f ms -> %.0f GB/s write" % (ms, 4.295/ms*1e3))
ms = t(lambda: x.fill_(1.5)); print("fill 4GiB: %.3f ms -> %.0f GB/s write" % (ms, 4.295/ms*1e3))
ms = t(lambda: y.copy_(x)); print("copy 4GiB: %.3f ms -> %.0f GB/s r+w" % (ms, 2*4.295/ms*1e3))
ms = t(lambda: x.sum()); print("sum 4GiB: %.3f ms -> %.0f GB/s read" % (ms, 4.295/ms*1e3))
