#!/usr/bin/env python
"""Times fix_size on the device (b200_resize_u8, 256x3 planes 518x518 -> 520x520) and fix_size + SWT with CUDA events."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402

args = argparse.Namespace(steps=10, warmup=3)
for level, wv in ((2, "haar"), (3, "sym4")):
    print(bench.bench_fix_size((256, 3, 518, 518), level, wv, args, torch.device("cuda", 0)), flush=True)
