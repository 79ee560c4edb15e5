"""Per-rank work of a multi-GPU step on ONE GPU: a slice of the queries against the whole database (what rank r of an
N-GPU HammingMapEngine evaluates), stage by stage, under the planner knobs given in the environment.
    python tools/slice_probe.py c3 628 [label]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from image_retrieval_wavelet_b200.engine.map_engine import HammingMapEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    qs = int(sys.argv[2]) if len(sys.argv) > 2 else 628
    q, ql, r, rl, k = bench.make_problem(name)
    dev = torch.device("cuda")
    q, ql, r, rl = q[:qs].contiguous().to(dev), ql[:qs].contiguous().to(dev), r.to(dev), rl.to(dev)
    flush = bench.l2_flusher(dev)
    eng = HammingMapEngine()
    out = eng.evaluate(q, ql, r, rl, k)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.evaluate(q, ql, r, rl, k, details=False)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    acc = {}
    for _ in range(5):
        flush()
        for kk, v in eng.stage_ms().items():
            acc[kk] = acc.get(kk, 0.0) + v / 5
    knobs = {k_: v for k_, v in os.environ.items() if k_.startswith("B200_")}
    print(f"{name} Q={qs} {knobs} step {sorted(ts)[len(ts)//2]:.4f} ms map={float(out[0]):.9f}",
          {kk: round(v, 4) for kk, v in acc.items() if "gated" not in kk and "round1" not in kk}, eng.plan_info(), flush=True)


if __name__ == "__main__":
    main()
