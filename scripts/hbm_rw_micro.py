#!/usr/bin/env python
"""Pure-write / pure-read / copy HBM bandwidth of this box (torch fill_, sum, copy_ on 1.5 GB buffers, CUDA events,
10 reps after 3 warm-ups): the denominators for store-dominated kernels (the stems write 10x what they read)."""
import torch


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def main():
    n = 768 * 1024 * 1024  # fp16 elements = 1.5 GB
    x = torch.empty(n, dtype=torch.float16, device="cuda")
    y = torch.empty(n, dtype=torch.float16, device="cuda")
    gb = n * 2 / 1e9
    w = t(lambda: x.fill_(1.0))
    z = t(lambda: x.zero_())
    r = t(lambda: x.view(torch.int32).sum())
    c = t(lambda: y.copy_(x))
    print("write (fill_)  %.0f GB/s ; write (zero_/memset) %.0f GB/s ; read (sum) %.0f GB/s ; copy %.0f GB/s (read + write)"
          % (gb / w, gb / z, gb / r, 2 * gb / c))


if __name__ == "__main__":
    main()
