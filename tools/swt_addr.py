import os, sys
sys.path.insert(0, "/root/repo")
import torch
from image_retrieval_wavelet_b200.transforms import swt2
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
shape = (256, 3, 520, 520)
x = torch.randint(0, 256, shape, dtype=torch.uint8).cuda()
for pad in (0, 1 << 20, 3 << 20, 17 << 20, 64 << 20, 1 << 30, (1 << 30) + (5 << 20), 2 << 30):
    dummy = torch.empty(max(pad, 1), dtype=torch.uint8, device="cuda")
    out = torch.empty(shape[:2] + (4,) + shape[2:], dtype=torch.float32, device="cuda")
    ts = []
    for r in range(4):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); swt2(x, "haar", 2, out=out); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"pad {pad>>20:5d} MiB  out ptr % 2MiB = {out.data_ptr() % (2<<20):8d}  x ptr % 2MiB = {x.data_ptr() % (2<<20):8d}  best {min(ts[1:])*1e3:.1f} us", flush=True)
    del out, dummy
    torch.cuda.empty_cache()
