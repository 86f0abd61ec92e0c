"""What connects the GPUs of this box?  Device-to-device copy bandwidth (copy engine, both directions) and a kernel
writing into the peer's memory, at the size of one rank's share of the walk matrix."""
import time
import torch

n = torch.cuda.device_count()
print("devices", n)
for a in range(min(n, 2)):
    for b in range(n):
        if a != b:
            print(a, b, "can_access_peer", torch.cuda.can_device_access_peer(a, b))
            break
if n >= 2:
    nbytes = 164 << 20
    x = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    y = torch.empty(nbytes, dtype=torch.uint8, device="cuda:1")
    for _ in range(3):
        y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    for label, dst, src in (("0->1", y, x), ("1->0", x, y)):
        dev = src.device
        with torch.cuda.device(dev):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / 10
        print(f"copy {label}: {ms:.3f} ms for {nbytes >> 20} MiB = {nbytes / ms / 1e6:.1f} GB/s")
