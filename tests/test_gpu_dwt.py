"""SURVEY.md §8 f4 on a B200: DWTTransform (pywt.wavedec2, mode 'symmetric', coarsest level) through b200_dwt2_fwd against
the oracle restatement.  Tolerance as for the SWT: per band, max |diff| <= 1e-5 * max |reference band| in float32 (with an
absolute floor of 1e-7 for bands that vanish identically)."""
import numpy as np
import pytest
import torch

from oracle import filters, swt_ref

pytestmark = pytest.mark.gpu

CASES = [
    ((8, 3, 32, 32), "haar", 1, np.uint8),         # config/transform/cifar_dwt.yaml shape
    ((8, 3, 32, 32), "haar", 2, np.uint8),
    ((4, 3, 32, 32), "db2", 2, np.float32),
    ((2, 3, 224, 224), "db4", 3, np.uint8),
    ((2, 1, 33, 47), "sym4", 2, np.float32),       # odd sizes: (N + F - 1) // 2
    ((1, 2, 64, 48), "bior4.4", 1, np.uint8),
    ((1, 1, 5, 3), "db4", 1, np.float32),          # shorter than the filter: repeated reflection
    ((1, 1, 1, 1), "haar", 1, np.uint8),
    ((3, 1, 40, 40), "coif1", 4, np.float32),
]


@pytest.mark.parametrize("shape,name,level,dtype", CASES)
def test_dwt2_matches_oracle(shape, name, level, dtype):
    from image_retrieval_wavelet_b200.transforms import dwt2

    rng = np.random.default_rng(sum(shape) + level)
    x = rng.integers(0, 256, shape).astype(np.uint8) if dtype == np.uint8 else rng.random(shape, dtype=np.float32)
    xf = x.astype(np.float32) / np.float32(255.0) if dtype == np.uint8 else x
    ref = swt_ref.dwt2_ref(xf, name, level)
    out = dwt2(torch.from_numpy(x).cuda(), name, level).cpu().numpy()
    assert out.shape == ref.shape and out.dtype == np.float32
    for band in range(4):
        # a detail band that cancels exactly in the reference (1 x 1 image) is only zero up to one float32 rounding of O(1) inputs
        tol = max(1e-5 * np.abs(ref[..., band, :, :]).max(), 1e-7)
        assert np.abs(out[..., band, :, :] - ref[..., band, :, :]).max() <= tol, (band, shape, name, level)


def test_dwt_transform_mirror():
    """DWTTransform(level, wavelet)(PIL image): fix_size, /255, wavedec2 per channel, [3, 4, H/2^L, W/2^L]."""
    Image = pytest.importorskip("PIL.Image")
    from image_retrieval_wavelet_b200.transforms import DWTTransform

    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (32, 32, 3), dtype=np.uint8)
    t = DWTTransform(level=2, wavelet="haar")
    assert repr(t) == "DWTTransform(shape='C,S,H/4,W/4', wavelet=haar, level=2)"
    out = t(Image.fromarray(img))
    assert tuple(out.shape) == (3, 4, 8, 8) and out.dtype == torch.float32 and not out.is_cuda
    ref = swt_ref.dwt2_ref(img.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0), "haar", 2)
    assert np.abs(out.numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    # a 30 x 30 image is resized to 32 x 32 first (fix_size), on the device for a uint8 batch
    small = rng.integers(0, 256, (2, 3, 30, 30), dtype=np.uint8)
    batch = t.forward(torch.from_numpy(small).cuda())
    assert tuple(batch.shape) == (2, 3, 4, 8, 8)
    single = t(Image.fromarray(np.ascontiguousarray(small[0].transpose(1, 2, 0))))
    assert torch.equal(single, batch[0].cpu())
    with pytest.raises(TypeError):
        t.forward(torch.zeros((1, 3, 8, 8), dtype=torch.float64, device="cuda"))
