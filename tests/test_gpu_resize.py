"""SURVEY.md §8 f3 on a B200: b200_resize_u8 (fix_size on the device) bit for bit against Pillow's own outputs (goldens and,
when importable, the installed library) and the oracle restatement; SWTTransform.forward applies it to uint8 batches."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, filters, resize_ref

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def resize_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "resize_golden.npz"))


def test_resize_matches_pillow_goldens(resize_golden):
    from image_retrieval_wavelet_b200.transforms import resize_u8

    g = resize_golden
    for name in g["cases"]:
        x, want = g[f"{name}/in"], g[f"{name}/out"]
        got = resize_u8(torch.from_numpy(x).cuda(), want.shape[-2:], str(g[f"{name}/filter"]))
        assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), want), name


@pytest.mark.parametrize("shape,size,filt", [((4, 3, 518, 518), (520, 520), "bicubic"), ((2, 3, 500, 375), (504, 376), "bicubic"),
                                             ((3, 1, 333, 500), (256, 384), "bilinear"), ((1, 2, 64, 48), (224, 224), "bilinear"),
                                             ((2, 1, 300, 301), (300, 304), "bicubic"), ((1, 1, 1, 1), (8, 8), "bicubic"),
                                             ((2, 2, 40, 40), (40, 40), "bicubic"), ((1, 1, 900, 700), (64, 50), "bicubic")])
def test_resize_matches_oracle(shape, size, filt):
    from image_retrieval_wavelet_b200.transforms import resize_u8

    x = np.random.default_rng(sum(shape) + size[0]).integers(0, 256, shape, dtype=np.uint8)
    x[..., ::5, ::3] = 255
    got = resize_u8(torch.from_numpy(x).cuda(), size, filt).cpu().numpy()
    assert np.array_equal(got, resize_ref.resize_ref(x, size, filt))


def test_resize_matches_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    from image_retrieval_wavelet_b200.transforms import resize_u8

    x = np.random.default_rng(1).integers(0, 256, (518, 518, 3), dtype=np.uint8)
    want = np.array(Image.fromarray(x).resize((520, 520), resample=Image.BICUBIC)).transpose(2, 0, 1)
    got = resize_u8(torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))).cuda(), (520, 520)).cpu().numpy()
    assert np.array_equal(got, want)


def test_forward_applies_fix_size_like_the_reference_call():
    """SWTTransform(level=3)(PIL 518x518) resizes to 520x520 first (custom_transforms.py:146); the batched device entry does
    the same to a uint8 batch: identical bits to the per-image path and parity with the oracle on the resized planes."""
    Image = pytest.importorskip("PIL.Image")
    from image_retrieval_wavelet_b200.transforms import SWTTransform

    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (2, 74, 70, 3), dtype=np.uint8)
    t = SWTTransform(level=3, wavelet="db2")
    batch = torch.from_numpy(np.ascontiguousarray(imgs.transpose(0, 3, 1, 2))).cuda()
    out = t.forward(batch)
    assert tuple(out.shape) == (2, 3, 4, 80, 72)
    lo, hi = filters.filter_bank("db2")
    for i in range(2):
        single = t(Image.fromarray(imgs[i]))                                   # host fix_size (PIL) + device SWT
        assert torch.equal(single, out[i].cpu())
    resized = resize_ref.fix_size_ref(imgs.transpose(0, 3, 1, 2), 3)
    ref = c_oracle.swt2(resized, lo, hi, 3)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    with pytest.raises(ValueError):
        t.forward(batch.float() / 255.0)                                       # float32 batches are not resized


def test_raw_stack_forward_applies_fix_size_too():
    Image = pytest.importorskip("PIL.Image")
    from image_retrieval_wavelet_b200.transforms import RawStackTransform

    img = np.random.default_rng(4).integers(0, 256, (30, 34, 3), dtype=np.uint8)
    t = RawStackTransform(level=2, copies=4)
    batch = t.forward(torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))[None].cuda())
    assert tuple(batch.shape) == (1, 3, 4, 32, 36)
    assert torch.equal(t(Image.fromarray(img)), batch[0].cpu())
    want = np.array(Image.fromarray(img).resize((36, 32), resample=Image.BICUBIC)).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    assert np.array_equal(batch[0, :, 2].cpu().numpy(), want)


def test_resize_rejects_bad_arguments():
    from image_retrieval_wavelet_b200.transforms import resize_u8

    x = torch.zeros((1, 8, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        resize_u8(x, (0, 8))
    with pytest.raises(ValueError):
        resize_u8(x, (8, 8), "lanczos")
    with pytest.raises(TypeError):
        resize_u8(x.float(), (8, 8))
    with pytest.raises(NotImplementedError):
        resize_u8(torch.zeros((1, 4000, 8), dtype=torch.uint8, device="cuda"), (16, 8))      # window wider than 64 taps
