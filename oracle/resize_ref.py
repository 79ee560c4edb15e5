"""CPU oracle for the pixel step in front of the SWT (test infrastructure, see ``oracle/__init__.py``).

Restates what ``BaseWaveletTransform.fix_size`` (``/root/reference/main/transforms/custom_transforms.py:132-139``) gets from
``PIL.Image.resize(size, resample=Image.BICUBIC)`` on an 8-bit image.  The arithmetic lives in Pillow (third-party; the
reference pins ``pillow==8.2.0``, requirements.txt:5; this container has Pillow 12.2.0 — the 8-bit resampler in
``src/libImaging/Resample.c`` is the same in both): ``precompute_coeffs`` (double-precision window weights, normalised),
``normalize_coeffs_8bpc`` (22-bit fixed point), ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc``
through a uint8 intermediate, a pass being skipped when its size does not change.

**Pinned**: ``tests/golden/make_golden_resize.py`` records Pillow's own outputs (``tests/golden/resize_golden.npz``) and
``tests/test_oracle_resize.py`` checks this restatement against them bit for bit — and against the installed Pillow
directly when it is importable.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


FILTERS = {"bicubic": (_bicubic, 2.0), "bilinear": (_bilinear, 1.0)}


def coeffs_ref(in_size, out_size, resample="bicubic"):
    """Resample.c ``precompute_coeffs`` + ``normalize_coeffs_8bpc``: (xmin [out], count [out], kk int32 [out, ksize])."""
    filt, fsupport = FILTERS[resample]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmins = np.zeros(out_size, dtype=np.int64)
    counts = np.zeros(out_size, dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = (k[x] / ww if ww != 0.0 else k[x]) * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if v < 0 else int(0.5 + v)          # C (int) truncates toward zero, like int()
        xmins[xx], counts[xx] = xmin, xmax
    return xmins, counts, kk


def _pass(img, out_size, axis, resample):
    """One 8-bit resampling pass along ``axis`` of ``[..., H, W]``."""
    xmins, counts, kk = coeffs_ref(img.shape[axis], out_size, resample)
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (out_size,), dtype=np.uint8)
    for xx in range(out_size):
        n, x0 = int(counts[xx]), int(xmins[xx])
        acc = (1 << (PRECISION_BITS - 1)) + (src[..., x0:x0 + n] * kk[xx, :n]).sum(-1)
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def resize_ref(img, size, resample="bicubic"):
    """``[..., H, W]`` uint8 -> ``[..., size[0], size[1]]`` uint8, Pillow's ``Image.resize`` per plane."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim >= 2
    ho, wo = int(size[0]), int(size[1])
    out = img
    if wo != img.shape[-1]:
        out = _pass(out, wo, -1, resample)          # horizontal first (ImagingResampleInner)
    if ho != img.shape[-2]:
        out = _pass(out, ho, -2, resample)
    return out.copy() if out is img else out


def fix_size_ref(img, level):
    """custom_transforms.py:132-139 on ``[..., H, W]`` uint8 planes."""
    factor = 2 ** level
    h, w = img.shape[-2:]
    new_h, new_w = int(np.ceil(h / factor) * factor), int(np.ceil(w / factor) * factor)
    if (new_h, new_w) == (h, w):
        return np.asarray(img)
    return resize_ref(img, (new_h, new_w), "bicubic")
