#!/usr/bin/env python
"""Generate ``tests/golden/swt_golden.npz`` from the REAL PyWavelets — the library whose ``pywt.swt2`` / ``pywt.wavedec2``
the reference calls (``/root/reference/main/transforms/custom_transforms.py:164,198``).

PyWavelets is installed neither in the build container nor on the GPU boxes of this project (``import pywt`` fails in
both; no wheel in /opt/wheelhouse), so this script has not produced a fixture yet and the SWT / DWT oracle stays
"parity unpinned" (DESIGN.md §3).  Run it on ANY machine that has ``pywt`` + numpy:

    python tests/golden/make_golden_swt.py            # writes tests/golden/swt_golden.npz

and commit the file: ``tests/test_oracle_swt_pywt.py`` then pins ``oracle.swt_ref`` (and, through it, the CUDA kernels)
to PyWavelets' own outputs on the C1 / C4 shapes and on every embedded filter bank.  The same test also compares against
a live ``pywt`` when one is importable, fixture or not.
"""
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "swt_golden.npz")
# (wavelet, level, H, W): C1, the C4 grid on small planes with the same divisibility, the study wavelets, odd aspect
SWT_CASES = [("haar", 1, 224, 224)] + [(w, l, 64, 72) for w in ("haar", "db2", "db4", "sym4") for l in (1, 2, 3)] + \
            [("bior4.4", 1, 48, 40), ("bior4.4", 2, 48, 40), ("db7", 2, 64, 64), ("haar", 4, 32, 48), ("db3", 1, 16, 24)]
DWT_CASES = [("haar", 2, 32, 32), ("db4", 3, 224, 224), ("db2", 1, 33, 47), ("sym4", 2, 50, 61), ("bior4.4", 2, 40, 40)]
FILTER_BANKS = ["haar", "db1", "db2", "db3", "db4", "db5", "db6", "db7", "db8", "sym4", "bior4.4"]


def inputs(h, w, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, (h, w), dtype=np.uint8).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def main():
    import pywt

    out = {"pywt_version": np.array(pywt.__version__)}
    for name in FILTER_BANKS:
        wv = pywt.Wavelet(name)
        out[f"filters/{name}/dec_lo"] = np.asarray(wv.dec_lo, np.float64)
        out[f"filters/{name}/dec_hi"] = np.asarray(wv.dec_hi, np.float64)
    for i, (name, level, h, w) in enumerate(SWT_CASES):
        x = inputs(h, w, 100 + i)
        ca, (ch, cv, cd) = pywt.swt2(x, name, level=level)[0]          # custom_transforms.py:164-165: coarsest level
        out[f"swt/{i}/x"] = x
        out[f"swt/{i}/bands"] = np.stack([ca, ch, cv, cd]).astype(np.float32)
        out[f"swt/{i}/dtype"] = np.array(str(np.asarray(ca).dtype))
    for i, (name, level, h, w) in enumerate(DWT_CASES):
        x = inputs(h, w, 200 + i)
        coeffs = pywt.wavedec2(x, name, level=level)                   # custom_transforms.py:198-199, mode 'symmetric'
        ca, (ch, cv, cd) = coeffs[0], coeffs[1]
        out[f"dwt/{i}/x"] = x
        out[f"dwt/{i}/bands"] = np.stack([ca, ch, cv, cd]).astype(np.float32)
    out["swt_cases"] = np.array([f"{n}|{l}|{h}|{w}" for n, l, h, w in SWT_CASES])
    out["dwt_cases"] = np.array([f"{n}|{l}|{h}|{w}" for n, l, h, w in DWT_CASES])
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} with PyWavelets {pywt.__version__}: {len(SWT_CASES)} swt2 cases, {len(DWT_CASES)} wavedec2 cases")


if __name__ == "__main__":
    sys.exit(main())
