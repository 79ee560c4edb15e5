"""CPU oracle for HP-SWT (test infrastructure; see ``oracle/__init__.py``).

Restates, for the path ``SWTTransform.__call__`` of the reference
(``/root/reference/main/transforms/custom_transforms.py:145-166``):

* ``fix_size``                     custom_transforms.py:132-139
* ``np.array(img).astype(float32) / 255``   custom_transforms.py:147
* ``pywt.swt2(ch, wavelet, level)`` and ``coeffs[0]``   custom_transforms.py:164-165
* the ``[3, 4, H, W]`` stacking     custom_transforms.py:149-157,166

``pywt`` (PyWavelets, unpinned in the reference's requirements.txt) is absent, so
``swt2`` is restated from its published algorithm:

* ``pywt.swt2`` = ``swtn`` over axes (-2, -1): for each level ``i`` (dilation
  ``2**i``), filter along axis -2 first then along axis -1; keys ``aa``=cA,
  ``da``=cH, ``ad``=cV, ``dd``=cD (first letter = axis -2); the next level
  consumes ``aa``; the returned list is coarsest-first.
* the 1-D step is PyWavelets' ``swt_a`` / ``swt_d``: the filter is upsampled by
  ``2**(level-1)`` and applied by ``downsampling_convolution_periodization`` with
  step 1, i.e. output ``o`` is ``sum_j e_filter[j] * x[(F_e/2 + o - j) mod N]``.
  For the non-zero taps this is
  ``y[n] = sum_j h[j] * x[(n + 2**(l-1) * (F/2 - j)) mod N]``.
* float32 input stays float32 (pywt picks the float32 C routines).

PARITY UNPINNED versus PyWavelets itself; pinned only on the documentation
examples reproduced in ``tests/test_oracle_swt.py``.
"""
import numpy as np

from .filters import filter_bank


def swt_step_1d(a, h, level, axis, dtype=np.float32):
    """One undecimated, periodised analysis step along ``axis`` (vectorised).

    ``a``: ndarray; ``h``: decomposition filter (length F, even); ``level`` >= 1
    selects the dilation ``2**(level-1)``.  Accumulates taps in ascending ``j``
    in ``dtype`` like the C loop does.
    """
    a = np.asarray(a, dtype=dtype)
    h = np.asarray(h, dtype=dtype)
    f = h.shape[0]
    s = 1 << (level - 1)
    out = np.zeros_like(a)
    for j in range(f):
        if h[j] == 0:
            continue
        shift = s * (f // 2 - j)
        out = out + h[j] * np.roll(a, -shift, axis=axis)      # a[(n + shift) mod N]
    return out.astype(dtype, copy=False)


def swt_step_1d_loop(x, h, level, dtype=np.float64):
    """Literal scalar restatement of the periodised convolution with the
    upsampled filter (small inputs only; used to cross-check the vectorised
    step, including inputs shorter than the dilated filter)."""
    x = np.asarray(x, dtype=dtype)
    n = x.shape[0]
    f = len(h)
    step = 1 << (level - 1)
    fe = f * step
    e = np.zeros(fe, dtype=dtype)
    for i in range(f):
        e[i * step] = h[i]
    out = np.zeros(n, dtype=dtype)
    for o in range(n):
        i = fe // 2 + o
        acc = dtype(0)
        for j in range(fe):
            acc = dtype(acc + e[j] * x[(i - j) % n])
        out[o] = acc
    return out


def swt2_ref(x, wavelet="haar", level=1, dtype=np.float32, all_levels=False):
    """``pywt.swt2(x, wavelet, level)`` restated.  ``x``: [..., H, W].

    Returns ``coeffs[0]`` as an array ``[..., 4, H, W]`` in the band order the
    reference stacks (cA, cH, cV, cD) = (LL, LH, HL, HH), or the whole
    coarsest-first list when ``all_levels``.
    """
    lo, hi = wavelet if isinstance(wavelet, (tuple, list)) else filter_bank(wavelet)
    x = np.asarray(x, dtype=dtype)
    hgt, wid = x.shape[-2], x.shape[-1]
    if hgt % (1 << level) or wid % (1 << level):
        raise ValueError("swt2 needs H and W divisible by 2**level")
    out = []
    a = x
    for lv in range(1, level + 1):
        ra = swt_step_1d(a, lo, lv, axis=-2, dtype=dtype)     # 'a' along axis -2
        rd = swt_step_1d(a, hi, lv, axis=-2, dtype=dtype)     # 'd' along axis -2
        aa = swt_step_1d(ra, lo, lv, axis=-1, dtype=dtype)
        ad = swt_step_1d(ra, hi, lv, axis=-1, dtype=dtype)    # cV
        da = swt_step_1d(rd, lo, lv, axis=-1, dtype=dtype)    # cH
        dd = swt_step_1d(rd, hi, lv, axis=-1, dtype=dtype)
        out.append(np.stack([aa, da, ad, dd], axis=-3))       # [cA, cH, cV, cD]
        a = aa
    out.reverse()
    return out if all_levels else out[0]


def fixed_size(w, h, level):
    """custom_transforms.py:132-136 — (new_w, new_h)."""
    factor = 2 ** level
    return int(np.ceil(w / factor) * factor), int(np.ceil(h / factor) * factor)


def swt_transform_ref(img_u8_hwc, wavelet="haar", level=1, dtype=np.float32):
    """``SWTTransform(level, wavelet)(PIL image)`` for an image whose size already
    satisfies ``fix_size`` (uint8 ``[H, W, 3]``) -> float32 ``[3, 4, H, W]``."""
    img = np.asarray(img_u8_hwc)
    assert img.ndim == 3 and img.shape[2] == 3
    w2, h2 = fixed_size(img.shape[1], img.shape[0], level)
    if (w2, h2) != (img.shape[1], img.shape[0]):
        raise ValueError("resize the image first (fix_size is PIL bicubic, host side)")
    x = img.astype(np.float32) / np.float32(255.0)
    chans = [swt2_ref(x[:, :, c], wavelet, level, dtype=dtype) for c in range(3)]
    return np.stack(chans).astype(np.float32)


def raw_stack_ref(img_u8_hwc, copies=4):
    """``RawStackTransform`` (custom_transforms.py:172-188)."""
    x = np.asarray(img_u8_hwc).astype(np.float32) / np.float32(255.0)
    return np.stack([np.stack([x[:, :, c]] * copies) for c in range(3)])


# --------------------------------------------------------------------------- decimated DWT
def dwt_step_1d(a, h, axis, dtype=np.float32):
    """One decimated analysis step of PyWavelets' ``dwt`` in mode 'symmetric' along ``axis``:
    ``y[o] = sum_j h[j] * xe[2o + 1 - j]``, ``o < (N + F - 1) // 2``, ``xe`` = half-sample symmetric extension.
    Taps accumulate in ascending ``j`` in ``dtype`` like the C loop."""
    a = np.moveaxis(np.asarray(a, dtype=dtype), axis, -1)
    h = np.asarray(h, dtype=dtype)
    n, f = a.shape[-1], h.shape[0]
    nout = (n + f - 1) // 2
    out = np.zeros(a.shape[:-1] + (nout,), dtype=dtype)
    o = np.arange(nout)
    for j in range(f):
        idx = 2 * o + 1 - j
        while ((idx < 0) | (idx >= n)).any():
            idx = np.where(idx < 0, -1 - idx, idx)
            idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
        out = out + h[j] * a[..., idx]
    return np.moveaxis(out.astype(dtype, copy=False), -1, axis)


def dwt2_ref(x, wavelet="haar", level=1, dtype=np.float32):
    """``DWTTransform._apply_wavelet`` (custom_transforms.py:196-200): ``pywt.wavedec2(x, wavelet, level=level)`` restated,
    coarsest level only -> ``[..., 4, H_L, W_L]`` = (cA, cH, cV, cD).  dwt2 filters axis -2 first, then axis -1;
    ``da`` = cH, ``ad`` = cV (first letter = axis -2).  PARITY UNPINNED versus PyWavelets (absent), like ``swt2_ref``."""
    lo, hi = wavelet if isinstance(wavelet, (tuple, list)) else filter_bank(wavelet)
    a = np.asarray(x, dtype=dtype)
    bands = None
    for _ in range(level):
        ra, rd = dwt_step_1d(a, lo, -2, dtype), dwt_step_1d(a, hi, -2, dtype)
        aa, ad = dwt_step_1d(ra, lo, -1, dtype), dwt_step_1d(ra, hi, -1, dtype)
        da, dd = dwt_step_1d(rd, lo, -1, dtype), dwt_step_1d(rd, hi, -1, dtype)
        bands = np.stack([aa, da, ad, dd], axis=-3)
        a = aa
    return bands
