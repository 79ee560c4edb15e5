"""Times b200_knn_topk on the BASELINE configs[2] cosine shape with the fused path on / off; ncu-friendly (few launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from image_retrieval_wavelet_b200.engine.get_knn import knn_topk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = torch.Generator().manual_seed(0)
refs = torch.nn.functional.normalize(torch.randn(117000, 768, generator=g), dim=1).cuda()
qs = torch.nn.functional.normalize(torch.randn(5000, 768, generator=g), dim=1).cuda()
flush = bench.l2_flusher(torch.device("cuda"))
for fused in ("1", "0"):
    os.environ["B200_KNN_FUSED"] = fused
    knn_topk(refs, qs, k, "cosine")
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        sc, idx = knn_topk(refs, qs, k, "cosine")
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"knn 5000x117000x768 k={k} fused={fused}: min {min(ts):.3f} ms mean {sum(ts)/len(ts):.3f} ms  checksum {idx.sum().item()}", flush=True)
