"""The resize oracle (oracle/resize_ref.py, a restatement of Pillow's 8-bit resampler) against Pillow's own outputs:
the committed goldens (tests/golden/resize_golden.npz) and, when Pillow is importable, the installed library directly."""
import os

import numpy as np
import pytest

from oracle import resize_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def resize_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "resize_golden.npz"))


def test_oracle_matches_pillow_goldens_bit_for_bit(resize_golden):
    g = resize_golden
    assert len(g["cases"]) >= 9
    for name in g["cases"]:
        x, want = g[f"{name}/in"], g[f"{name}/out"]
        got = resize_ref.resize_ref(x, want.shape[-2:], str(g[f"{name}/filter"]))
        assert got.dtype == np.uint8 and np.array_equal(got, want), name


def test_oracle_matches_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for (h, w), level in [((37, 54), 2), ((50, 30), 3), ((64, 64), 2), ((518 // 7, 518 // 7), 3)]:
        x = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        factor = 2 ** level
        nh, nw = -(-h // factor) * factor, -(-w // factor) * factor
        want = np.array(Image.fromarray(x).resize((nw, nh), resample=Image.BICUBIC)) if (nh, nw) != (h, w) else x
        got = resize_ref.fix_size_ref(x.transpose(2, 0, 1), level)
        assert np.array_equal(got, want.transpose(2, 0, 1)), (h, w, level)
    x = rng.integers(0, 256, (120, 90), dtype=np.uint8)
    for size, filt, pil in [((256, 192), "bilinear", Image.BILINEAR), ((30, 40), "bilinear", Image.BILINEAR),
                            ((31, 17), "bicubic", Image.BICUBIC)]:
        want = np.array(Image.fromarray(x).resize((size[1], size[0]), resample=pil))
        assert np.array_equal(resize_ref.resize_ref(x, size, filt), want), (size, filt)


def test_known_answers():
    """Constant images stay constant (weights sum to 2^22 up to rounding that clip8 absorbs); identity size is a copy."""
    for v in (0, 1, 127, 254, 255):
        x = np.full((2, 9, 11), v, np.uint8)
        assert np.all(resize_ref.resize_ref(x, (12, 16)) == v) and np.all(resize_ref.resize_ref(x, (4, 5), "bilinear") == v)
    x = np.arange(35, dtype=np.uint8).reshape(5, 7)
    assert np.array_equal(resize_ref.resize_ref(x, (5, 7)), x)
    xm, cnt, kk = resize_ref.coeffs_ref(518, 520)
    assert kk.shape == (520, 5) and np.all(np.abs(kk.sum(1) - (1 << 22)) <= 2) and xm[0] == 0 and xm[-1] + cnt[-1] == 518
