#!/usr/bin/env python
"""Times b200_maphashing_host (pinned host float32 in, mAP out) on a bench workload: python tools/host_e2e.py c3 [reps]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from image_retrieval_wavelet_b200 import _cabi
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
desc, nq, n, bits, nlab, _, _ = bench.WORKLOADS[name]
q, ql, r, rl, k = bench.make_problem(name)
hq, hql, hr, hrl = (t.pin_memory() for t in (q, ql, r, rl))
lib = _cabi.load()
m, bad = ctypes.c_double(), ctypes.c_int()
def step():
    rc = lib.b200_maphashing_host(hq.data_ptr(), hql.data_ptr(), hr.data_ptr(), hrl.data_ptr(), nq, n, bits, nlab, 0, k, None, None,
                                  ctypes.addressof(m), ctypes.addressof(bad))
    assert rc == 0, rc
for _ in range(3): step()
t0 = time.perf_counter()
for _ in range(reps): step()
dt = (time.perf_counter() - t0) / reps
print(f"{name}: chunks={os.environ.get('B200_HOST_CHUNKS','default')} {dt*1e3:.3f} ms/step  mAP={m.value:.6f}", flush=True)
