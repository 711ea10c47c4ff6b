"""Pure-write, pure-read and copy bandwidth of HBM (torch kernels, CUDA events): is a write-only stream slower than a copy?"""
import torch
dev = torch.device("cuda:0")
n = 1 << 30                                  # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(it):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_());            print(f"write (zero_)    {4 * n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: a.fill_(1.5));         print(f"write (fill_)    {4 * n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: torch.sum(a));         print(f"read  (sum)      {4 * n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: b.copy_(a));           print(f"copy  (r+w)      {8 * n / ms / 1e6:8.1f} GB/s  ({4 * n / ms / 1e6:.1f} each way)")
ms = t(lambda: torch.add(a, 1.0, out=b)); print(f"add out= (r+w)   {8 * n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: torch.add(a, b, out=b));   print(f"a+b->b (2r+w)    {12 * n / ms / 1e6:8.1f} GB/s")
