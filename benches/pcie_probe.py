#!/usr/bin/env python
"""Ceiling of the end-to-end leg: plain pinned host-to-device copies of the cfg 2 stream (8.32 GB) in one piece and in
512 MiB / 64 MiB chunks, timed with CUDA events.  Needs a GPU.  Output kept in profiles/r02_pcie_probe.txt."""
import torch

n = 8_324_160_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 512 << 20, 64 << 20):
    for rep in range(3):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"pinned H2D {n / 1e9:.2f} GB in chunks of {chunk >> 20} MiB: {ms:.2f} ms = {n / ms / 1e6:.2f} GB/s")
