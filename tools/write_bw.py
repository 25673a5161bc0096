"""Developer probe: pure-write and copy bandwidth at the C2 output size (ceiling for the eval kernel)."""
import torch
dev = torch.device("cuda:0")
for mb in (128, 512):
    n = mb * 1024 * 1024 // 8
    bufs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)]
    src = torch.ones(n, dtype=torch.float64, device=dev)
    for name, fn in (("fill", lambda b: b.fill_(1.0)), ("copy", lambda b: b.copy_(src))):
        for i in range(8):
            fn(bufs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(100):
            fn(bufs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 10
        by = n * 8 * (2 if name == "copy" else 1)
        print(f"{name} {mb} MB: {us:.2f} us  {by / us / 1e3:.0f} GB/s")
