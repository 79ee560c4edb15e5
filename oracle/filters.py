"""Decomposition filter banks for the oracle (test infrastructure).

PyWavelets is absent, so the oracle carries its own filters:

* Daubechies ``dbN`` are *computed* here by spectral factorisation of the
  Daubechies half-band polynomial (minimum-phase root choice), i.e. from the
  published construction, independently of the product's embedded tables in
  ``image_retrieval_wavelet_b200/transforms/wavelets.py``.  ``tests`` compare
  the two.
* ``sym4``, ``coif1`` and the ``bior`` pairs are PyWavelets' published tables
  (they were re-verified numerically: orthonormality for the orthogonal ones,
  the half-band perfect-reconstruction identity for the biorthogonal pairs).

Convention (PyWavelets): ``dec_lo`` is the time-reversed scaling filter and, for
orthogonal wavelets, ``dec_hi[k] = (-1)**(k+1) * dec_lo[F-1-k]``.
"""
import math

import numpy as np

_SYM4 = [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161,
         0.29785779560527736, -0.09921954357684722, -0.012603967262037833, 0.0322231006040427]
_COIF1 = [-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554,
          0.3378976624578092, -0.0727326195128539]
_BIOR = {
    "bior1.1": ([0.7071067811865476, 0.7071067811865476], [-0.7071067811865476, 0.7071067811865476]),
    "bior1.3": ([-0.08838834764831845, 0.08838834764831845, 0.7071067811865476, 0.7071067811865476,
                 0.08838834764831845, -0.08838834764831845],
                [0.0, 0.0, -0.7071067811865476, 0.7071067811865476, 0.0, 0.0]),
    "bior2.2": ([0.0, -0.1767766952966369, 0.3535533905932738, 1.0606601717798214, 0.3535533905932738,
                 -0.1767766952966369],
                [0.0, 0.3535533905932738, -0.7071067811865476, 0.3535533905932738, 0.0, 0.0]),
    "bior4.4": ([0.0, 0.03782845550726404, -0.023849465019556843, -0.11062440441843718, 0.37740285561283066,
                 0.8526986790088938, 0.37740285561283066, -0.11062440441843718, -0.023849465019556843,
                 0.03782845550726404],
                [0.0, -0.06453888262869706, 0.04068941760916406, 0.41809227322161724, -0.7884856164055829,
                 0.41809227322161724, 0.04068941760916406, -0.06453888262869706, 0.0, 0.0]),
}


def daubechies_dec_lo(order):
    """dec_lo of dbN by spectral factorisation (Daubechies 1988 construction)."""
    n = int(order)
    if n < 1:
        raise ValueError("order must be >= 1")
    h = np.array([1.0 + 0j])
    for _ in range(n):
        h = np.convolve(h, [1.0, 1.0])
    if n > 1:
        # P(y) = sum_{k<N} C(N-1+k, k) y^k with y = (2 - z - 1/z) / 4
        coeffs = [math.comb(n - 1 + k, k) for k in range(n)]
        for y in np.roots(coeffs[::-1]):
            r = np.roots([1.0, -(2.0 - 4.0 * y), 1.0])
            h = np.convolve(h, [1.0, -r[np.argmin(np.abs(r))]])   # root inside the unit circle
    h = h.real
    h = h / h.sum() * math.sqrt(2.0)          # rec_lo
    return h[::-1].copy()                      # dec_lo = reversed rec_lo


def orthogonal_dec_hi(dec_lo):
    lo = np.asarray(dec_lo, dtype=np.float64)
    f = lo.shape[0]
    return np.array([(-1.0) ** (k + 1) * lo[f - 1 - k] for k in range(f)])


def filter_bank(name):
    """Return (dec_lo, dec_hi) float64 arrays for a PyWavelets wavelet name."""
    name = str(name).lower()
    if name in _BIOR:
        lo, hi = _BIOR[name]
        return np.array(lo), np.array(hi)
    if name == "haar":
        lo = daubechies_dec_lo(1)
    elif name.startswith("db") and name[2:].isdigit():
        lo = daubechies_dec_lo(int(name[2:]))
    elif name in ("sym2", "sym3"):
        lo = daubechies_dec_lo(int(name[3:]))
    elif name == "sym4":
        lo = np.array(_SYM4)
    elif name == "coif1":
        lo = np.array(_COIF1)
    else:
        raise ValueError(f"oracle has no filter bank for wavelet {name!r}")
    return lo, orthogonal_dec_hi(lo)
