#!/usr/bin/env python
"""Runs the SWT kernel on a fixed list of BASELINE shapes (profiling driver: `ncu ... python tools/swt_cases.py`).
Usage: python tools/swt_cases.py [reps]   — prints CUDA-event times; the product path only (no oracle)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from image_retrieval_wavelet_b200.transforms import swt2  # noqa: E402

CASES = [((64, 3, 224, 224), "haar", 1, torch.uint8), ((64, 3, 224, 224), "haar", 1, torch.float32),
         ((256, 3, 518, 518), "haar", 1, torch.uint8), ((256, 3, 518, 518), "db4", 1, torch.uint8),
         ((256, 3, 520, 520), "haar", 2, torch.uint8), ((256, 3, 520, 520), "db2", 3, torch.uint8),
         ((256, 3, 520, 520), "sym4", 3, torch.uint8), ((256, 3, 520, 520), "db2", 1, torch.uint8),
         ((256, 3, 520, 520), "sym4", 2, torch.uint8), ((256, 3, 518, 518), "bior4.4", 1, torch.uint8)]

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i, (shape, wv, lv, dt) in enumerate(CASES):
    if only and str(i) not in only:
        continue
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g).cuda()
    if dt == torch.float32:
        x = x.float() / 255.0
    out = torch.empty(shape[:2] + (4,) + shape[2:], dtype=torch.float32, device="cuda")
    ts = []
    for r in range(reps + 1):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        swt2(x, wv, lv, out=out)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    px = shape[0] * shape[1] * shape[2] * shape[3]
    by = px * ((1 if dt == torch.uint8 else 4) + 16)
    t = min(ts[1:]) if reps > 0 else ts[0]
    print(f"case {i}: {wv} L{lv} {shape} {str(dt).split('.')[-1]}: {t * 1e3:.1f} us  {by / t / 1e6:.0f} GB/s  frac {by / t / 1e6 / 6552.6:.3f}", flush=True)
