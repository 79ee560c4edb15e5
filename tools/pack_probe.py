"""Achieved HBM rate of the bit-packing kernels (b200_pack_codes / b200_pack_labels) on bench-shaped inputs.
    python tools/pack_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from image_retrieval_wavelet_b200 import _cabi  # noqa: E402
from image_retrieval_wavelet_b200.engine import hamming as H  # noqa: E402


def timed(fn, flush, reps=10):
    ts = []
    for _ in range(reps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda")
    flush = bench.l2_flusher(dev)
    peak = bench.measured_peaks()
    print("peaks:", peak)
    for rows, cols, kind in ((117000, 128, "codes"), (117000, 80, "labels"), (1000000, 64, "codes"), (1000000, 80, "labels"),
                             (125000, 64, "codes"), (14625, 128, "codes"), (5000, 128, "codes")):
        g = torch.Generator(device="cpu").manual_seed(1)
        x = (torch.rand(rows, cols, generator=g) < 0.5).float()
        if kind == "codes":
            x = x * 2 - 1
        x = x.to(dev)
        lib = _cabi.load()
        words = _cabi.code_words(cols) if kind == "codes" else _cabi.label_words(cols)
        out = torch.empty(((rows + 1) // 2 * 2, words), dtype=torch.int64, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        call = lib.b200_pack_codes if kind == "codes" else lib.b200_pack_labels
        fn = lambda: _cabi.check(call(_cabi.ptr(x), rows, cols, _cabi.ptr(out), _cabi.ptr(bad), _cabi.stream_ptr()), "pack")  # noqa: E731
        fn()
        # independent check: the same words from torch integer arithmetic
        bits = (x > 0).to(torch.int64) if kind == "codes" else (x != 0).to(torch.int64)
        padc = torch.zeros(rows, words * 64, dtype=torch.int64, device=dev)
        padc[:, :cols] = bits
        ref = (padc.view(rows, words, 64) << torch.arange(64, device=dev)).sum(-1)
        assert torch.equal(ref, out[:rows]) and int(bad.item()) == 0, "packed words differ from the reference"
        ms = timed(fn, flush)
        byts = rows * cols * 4 + rows * words * 8
        print(f"{kind} {rows}x{cols}: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s algorithmic ({byts/1e6:.1f} MB)", flush=True)


if __name__ == "__main__":
    main()
