#!/usr/bin/env python
"""One cosine k-NN call on the BASELINE configs[2] shape (profiling driver for ncu): python tools/knn_case.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_retrieval_wavelet_b200.engine.get_knn import knn_topk
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = torch.Generator().manual_seed(0)
r = torch.nn.functional.normalize(torch.randn(117000, 768, generator=g), dim=1).cuda()
q = torch.nn.functional.normalize(torch.randn(5000, 768, generator=g), dim=1).cuda()
for _ in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); knn_topk(r, q, 2048, "cosine"); e.record(); torch.cuda.synchronize()
    print(f"knn_topk 5000x117000x768 k=2048: {s.elapsed_time(e):.3f} ms", flush=True)
