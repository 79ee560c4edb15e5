"""Decomposition filter banks (PyWavelets convention) for the SWT kernel.

The reference resolves ``wavelet`` names through PyWavelets inside ``pywt.swt2``
(``/root/reference/main/transforms/custom_transforms.py:164``); PyWavelets is not a dependency here, so the tables its
configs use are embedded: ``haar`` (config/transform/basic_swt.yaml), ``db4`` and ``bior4.4``
(studies/mflickr_wavelet_type_ablation.yaml), plus the ``db2`` / ``sym4`` of BASELINE.json's ablation and a few more
short filters.  Values are PyWavelets' published tables; ``tests/test_wavelet_tables.py`` re-derives the Daubechies
ones by spectral factorisation and checks orthonormality / perfect reconstruction of all of them.

``dec_lo`` is the time-reversed scaling filter; for orthogonal wavelets ``dec_hi[k] = (-1)**(k+1) * dec_lo[F-1-k]``.
A custom bank can be passed as ``(dec_lo, dec_hi)``.
"""
import math

_S2 = 0.7071067811865476

_DEC_LO = {
    "haar": [_S2, _S2],
    "db2": [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    "db3": [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313, 0.8068915093133388,
            0.3326705529509569],
    "db4": [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
            -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523],
    "db5": [0.003335725285001549, -0.012580751999015526, -0.006241490213011705, 0.07757149384006515,
            -0.03224486958502952, -0.24229488706619015, 0.13842814590110342, 0.7243085284385744, 0.6038292697974729,
            0.160102397974125],
    "sym4": [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161, 0.29785779560527736,
             -0.09921954357684722, -0.012603967262037833, 0.0322231006040427],
    "coif1": [-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554, 0.3378976624578092,
              -0.0727326195128539],
}
_DEC_LO["db1"] = _DEC_LO["haar"]
_DEC_LO["sym2"] = _DEC_LO["db2"]
_DEC_LO["sym3"] = _DEC_LO["db3"]

_BIORTHOGONAL = {
    "bior1.1": ([_S2, _S2], [-_S2, _S2]),
    "bior1.3": ([-0.08838834764831845, 0.08838834764831845, _S2, _S2, 0.08838834764831845, -0.08838834764831845],
                [0.0, 0.0, -_S2, _S2, 0.0, 0.0]),
    "bior2.2": ([0.0, -0.1767766952966369, 0.3535533905932738, 1.0606601717798214, 0.3535533905932738,
                 -0.1767766952966369],
                [0.0, 0.3535533905932738, -_S2, 0.3535533905932738, 0.0, 0.0]),
    "bior4.4": ([0.0, 0.03782845550726404, -0.023849465019556843, -0.11062440441843718, 0.37740285561283066,
                 0.8526986790088938, 0.37740285561283066, -0.11062440441843718, -0.023849465019556843,
                 0.03782845550726404],
                [0.0, -0.06453888262869706, 0.04068941760916406, 0.41809227322161724, -0.7884856164055829,
                 0.41809227322161724, 0.04068941760916406, -0.06453888262869706, 0.0, 0.0]),
}

MAX_FILTER_LENGTH = 20


def wavelist():
    return sorted(list(_DEC_LO) + list(_BIORTHOGONAL))


def filter_bank(wavelet):
    """``(dec_lo, dec_hi)`` as lists of Python floats for a wavelet name or an explicit pair."""
    if isinstance(wavelet, (tuple, list)) and len(wavelet) == 2 and not isinstance(wavelet[0], (int, float)):
        lo, hi = [float(v) for v in wavelet[0]], [float(v) for v in wavelet[1]]
    elif hasattr(wavelet, "dec_lo") and hasattr(wavelet, "dec_hi"):        # a pywt.Wavelet-like object
        lo, hi = [float(v) for v in wavelet.dec_lo], [float(v) for v in wavelet.dec_hi]
    else:
        name = str(wavelet).lower()
        if name in _BIORTHOGONAL:
            lo, hi = _BIORTHOGONAL[name]
        elif name in _DEC_LO:
            lo = _DEC_LO[name]
            f = len(lo)
            hi = [(-1.0) ** (k + 1) * lo[f - 1 - k] for k in range(f)]
        else:
            # pywt raises ValueError("Unknown wavelet name ...") for names it does not know
            raise ValueError(f"Unknown wavelet name {wavelet!r}; embedded banks: {', '.join(wavelist())} "
                             "(pass (dec_lo, dec_hi) for any other)")
        lo, hi = list(lo), list(hi)
    if len(lo) != len(hi) or len(lo) < 2 or len(lo) % 2 or len(lo) > MAX_FILTER_LENGTH:
        raise ValueError("decomposition filters must have the same even length in [2, 20]")
    if not all(math.isfinite(v) for v in lo + hi):
        raise ValueError("decomposition filters must be finite")
    return lo, hi
