#!/usr/bin/env python
"""Generate ``tests/golden/resize_golden.npz``: outputs of the REAL Pillow resampler (``Image.resize``), the third-party
arithmetic behind ``BaseWaveletTransform.fix_size`` (/root/reference/main/transforms/custom_transforms.py:132-139) and
behind torchvision's ``Resize`` on PIL images in the reference's eval transforms.

    python tests/golden/make_golden_resize.py
"""
import os

import numpy as np
import PIL
from PIL import Image

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_golden.npz")
RESAMPLE = {"bicubic": Image.BICUBIC, "bilinear": Image.BILINEAR}

# (H, W) -> (H', W'), filter, mode
CASES = [
    ((37, 54), (40, 56), "bicubic", "RGB"),      # fix_size at level 3 (both axes grow)
    ((40, 54), (40, 56), "bicubic", "RGB"),      # only the width changes: vertical pass skipped
    ((37, 56), (40, 56), "bicubic", "L"),        # only the height changes
    ((130, 130), (132, 132), "bicubic", "RGB"),  # 518 -> 520 in miniature (scale just below 1)
    ((97, 131), (64, 64), "bicubic", "RGB"),     # reduction: antialiasing window wider than 2
    ((97, 131), (48, 80), "bilinear", "RGB"),    # torchvision Resize on PIL (bilinear, antialiased)
    ((33, 21), (99, 64), "bilinear", "L"),       # enlargement
    ((5, 7), (8, 8), "bicubic", "L"),            # windows clipped at both borders
    ((1, 9), (4, 12), "bicubic", "L"),
]


def main():
    out = {"pillow_version": np.array(PIL.__version__)}
    names = []
    rng = np.random.default_rng(2024)
    for i, ((h, w), (ho, wo), filt, mode) in enumerate(CASES):
        shape = (h, w, 3) if mode == "RGB" else (h, w)
        arr = rng.integers(0, 256, shape, dtype=np.uint8)
        if i % 3 == 0:                         # saturated edges exercise clip8
            arr[::2, ::3] = 255
            arr[1::2, 1::3] = 0
        res = np.array(Image.fromarray(arr, mode=mode).resize((wo, ho), resample=RESAMPLE[filt]))
        name = f"case{i}_{filt}_{h}x{w}_to_{ho}x{wo}_{mode}"
        planes = lambda a: a.transpose(2, 0, 1) if a.ndim == 3 else a[None]
        out[f"{name}/in"], out[f"{name}/out"] = np.ascontiguousarray(planes(arr)), np.ascontiguousarray(planes(res))
        out[f"{name}/filter"] = np.array(filt)
        names.append(name)
        print(name, res.shape)
    out["cases"] = np.array(names)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} bytes), Pillow {PIL.__version__}")


if __name__ == "__main__":
    main()
