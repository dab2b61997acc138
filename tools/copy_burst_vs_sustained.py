#!/usr/bin/env python
"""The copy the HBM peak of MEASURED_PEAKS.json is defined by (torch b.copy_(a), 1 Gi bf16 elements = 2 GiB read + 2 GiB
written), timed alone after an idle pause and inside a run of back-to-back copies."""
import time
import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
a.fill_(1.0); b.copy_(a); torch.cuda.synchronize()
nbytes = 2 * a.numel() * a.element_size()


def timed(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for rep in range(3):
    time.sleep(3.0)
    alone = [timed(1) for _ in range(3)]
    run = timed(200)   # ~130 ms of continuous traffic, as bench.py's timed region
    run2 = timed(1000)
    print(f"copy alone after a 3 s pause: {nbytes / min(alone) * 1e-6:.0f} GB/s; 200 back to back: {nbytes / run * 1e-6:.0f} GB/s; "
          f"1000 back to back: {nbytes / run2 * 1e-6:.0f} GB/s", flush=True)
