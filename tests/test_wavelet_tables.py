"""The embedded filter tables of the product against the oracle's independently derived ones."""
import numpy as np
import pytest

from image_retrieval_wavelet_b200.transforms import wavelets
from oracle import filters


@pytest.mark.parametrize("name", ["haar", "db1", "db2", "db3", "db4", "db5", "sym2", "sym3", "sym4", "coif1", "bior1.1", "bior1.3",
                                  "bior2.2", "bior4.4"])
def test_tables_agree_with_oracle(name):
    lo, hi = wavelets.filter_bank(name)
    olo, ohi = filters.filter_bank(name)
    assert len(lo) == len(olo) and np.abs(np.array(lo) - olo).max() < 1e-10 and np.abs(np.array(hi) - ohi).max() < 1e-10
    # identical after rounding to the float32 the kernels use
    assert np.array_equal(np.array(lo, np.float32), olo.astype(np.float32))


def test_custom_bank_and_errors():
    lo, hi = wavelets.filter_bank(([0.5, 0.5], [-0.5, 0.5]))
    assert lo == [0.5, 0.5] and hi == [-0.5, 0.5]

    class W:
        dec_lo, dec_hi = [1.0, 0.0], [0.0, 1.0]
    assert wavelets.filter_bank(W()) == ([1.0, 0.0], [0.0, 1.0])
    with pytest.raises(ValueError):
        wavelets.filter_bank("no_such_wavelet")
    with pytest.raises(ValueError):
        wavelets.filter_bank(([1.0, 2.0, 3.0], [1.0, 2.0, 3.0]))
    assert "haar" in wavelets.wavelist() and "bior4.4" in wavelets.wavelist()
