"""Stage times of the select pipeline on a bench workload under B200_STC_DBG probe modes (tensor-core select kernel).
    python tools/stc_probe.py [c3] [modes, comma separated]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from image_retrieval_wavelet_b200.engine.map_engine import HammingMapEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    modes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,3,8,16").split(",")]
    q, ql, r, rl, k = bench.make_problem(name)
    dev = torch.device("cuda")
    q, ql, r, rl = q.to(dev), ql.to(dev), r.to(dev), rl.to(dev)
    eng = HammingMapEngine()
    for m in modes:
        os.environ["B200_STC_DBG"] = str(m)
        out = eng.evaluate(q, ql, r, rl, k)
        torch.cuda.synchronize()
        acc = {}
        for _ in range(5):
            for kk, v in eng.stage_ms().items():
                acc[kk] = acc.get(kk, 0.0) + v / 5
        print(f"{name} dbg={m} map={float(out[0]):.9f}", {kk: round(v, 4) for kk, v in acc.items() if "gated" not in kk and "round1" not in kk}, flush=True)


if __name__ == "__main__":
    main()
