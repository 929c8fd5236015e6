"""Probe: pure-write, pure-read and copy bandwidth on this GPU for BEV-sized buffers (CUDA events)."""
import torch
dev = "cuda:0"
def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3  # us
for mb in (82, 164, 656, 2048):
    n = mb * 1000 * 1000 // 4
    bufs = [torch.empty(n, device=dev) for _ in range(max(1, 700 // mb))]   # rotate so that the set exceeds L2
    src = [torch.randn(n, device=dev) for _ in range(len(bufs))]
    k = [0]
    def fill():
        bufs[k[0] % len(bufs)].zero_(); k[0] += 1
    def rd():
        src[k[0] % len(src)].sum(); k[0] += 1
    def cp():
        bufs[k[0] % len(bufs)].copy_(src[k[0] % len(src)]); k[0] += 1
    tf, tr, tc = timeit(fill), timeit(rd), timeit(cp)
    print("%5d MB x%d: fill %.1f us (%.2f TB/s)  read(sum) %.1f us (%.2f TB/s)  copy %.1f us (%.2f TB/s r+w)" % (
        mb, len(bufs), tf, mb / tf, tr, mb / tr, tc, 2 * mb / tc))
