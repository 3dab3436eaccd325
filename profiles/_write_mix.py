"""Context for the write-heavy streaming kernels (normalize u8 -> fp32: 1 byte read, 4 written): what plain library kernels reach on
this GPU for pure writes and for the same 1:4 mix, next to the copy peak the rooflines are quoted against."""
import torch

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n

n = 128 * 3840 * 2160 * 3
src = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
dst = torch.empty(n, dtype=torch.float32, device="cuda")
a = torch.empty(n // 2, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
ms = t(lambda: b.copy_(a)); print(f"copy fp32 (1:1)            {ms:8.3f} ms  {2 * a.numel() * 4 / ms / 1e6:8.1f} GB/s")
ms = t(lambda: dst.fill_(1.0)); print(f"fill fp32 (pure write)     {ms:8.3f} ms  {n * 4 / ms / 1e6:8.1f} GB/s")
ms = t(lambda: torch.cuda.current_stream().synchronize() or dst.zero_()); print(f"zero fp32 (memset)         {ms:8.3f} ms  {n * 4 / ms / 1e6:8.1f} GB/s")
ms = t(lambda: dst.copy_(src)); print(f"u8 -> fp32 convert (1:4)   {ms:8.3f} ms  {n * 5 / ms / 1e6:8.1f} GB/s")
