import os, sys
sys.path.insert(0, '/root/repo')
import torch, bench
from image_retrieval_wavelet_b200.engine.map_engine import HammingMapEngine
q, ql, r, rl, k = bench.make_problem("c3")
dev = torch.device("cuda")
q, ql, r, rl = q.to(dev), ql.to(dev), r.to(dev), rl.to(dev)
eng = HammingMapEngine()
out = eng.evaluate(q, ql, r, rl, k)
torch.cuda.synchronize()
for i in range(4):
    st = eng.stage_ms()
    print(i, {kk: round(v, 4) for kk, v in st.items()}, flush=True)
from image_retrieval_wavelet_b200.engine import hamming as H
qc, rc = H.pack_codes(q), H.pack_codes(r)
qlp, rlp = H.pack_labels(ql), H.pack_labels(rl)
import time
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m, ap, ts = H.hamming_map(qc, qlp, rc, rlp, k)
    torch.cuda.synchronize(); print('hamming_map full', round((time.perf_counter() - t0) * 1e3, 3), 'ms', float(m))
