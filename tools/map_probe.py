"""Times b200_hamming_map (packed inputs resident) on a bench workload with the select pipeline on / off.
    python tools/map_probe.py [c3|c5|...] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from image_retrieval_wavelet_b200.engine import hamming as H  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    q, ql, r, rl, k = bench.make_problem(name)
    dev = torch.device("cuda")
    qc, rc = H.pack_codes(q.to(dev), on_nonbinary="sign"), H.pack_codes(r.to(dev), on_nonbinary="sign")
    qlp, rlp = H.pack_labels(ql.to(dev)), H.pack_labels(rl.to(dev))
    flush = bench.l2_flusher(dev)
    results = {}
    for sel in ("1", "0"):
        os.environ["B200_MAP_SELECT"] = sel
        m, ap, ts, ws = H.hamming_map(qc, qlp, rc, rlp, k, return_workspace=True)
        torch.cuda.synchronize()
        st = H.select_status(ws)
        times = []
        for _ in range(reps):
            flush()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            H.hamming_map(qc, qlp, rc, rlp, k, workspace=ws)
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        results[sel] = (m.item(), int(ts.sum().item()), min(times), sum(times) / len(times))
        print(f"{name} select={sel}: map={m.item():.9f} hits={int(ts.sum().item())} min {min(times):.3f} ms mean {sum(times)/len(times):.3f} ms "
              f"status={st} plan: S={ws.plan.S} seg={ws.plan.seg_len} selS={ws.plan.sel_S} selseg={ws.plan.sel_seg_len} "
              f"ws={ws.plan.workspace_bytes/2**20:.0f} MiB", flush=True)
    assert results["1"][:2] == results["0"][:2], "select and three-stage pipelines disagree"


if __name__ == "__main__":
    main()
