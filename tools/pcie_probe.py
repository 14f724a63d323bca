"""PCIe ceiling of the end-to-end step: the step's host->device bytes (1.024 GB) and device->host bytes (0.256 GB)
from pinned memory, alone and concurrently on two streams, timed with CUDA events."""
import torch

dev = torch.device("cuda:0")
up_h = torch.empty(1_024_000_000, dtype=torch.uint8).pin_memory()
dn_h = torch.empty(256_000_000, dtype=torch.uint8).pin_memory()
up_d = torch.empty_like(up_h, device=dev)
dn_d = torch.empty_like(dn_h, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    s1.synchronize()
    s2.synchronize()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def up():
    with torch.cuda.stream(s1):
        up_d.copy_(up_h, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        dn_h.copy_(dn_d, non_blocking=True)


def both():
    up()
    down()


import time
for name, fn, nbytes in (("H2D 1.024 GB alone", up, 1.024e9), ("D2H 0.256 GB alone", down, 0.256e9), ("both, two streams", both, 1.28e9)):
    # wall clock: the copies run on side streams
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 8 * 1e3
    print(f"{name}: {ms:.2f} ms per step  ({nbytes / ms / 1e6:.1f} GB/s aggregate)")
