"""B200-native (sm_100a) implementation of the two data-parallel hot paths of ArseneAmoya/image-retrieval-wavelet:

* ``transforms`` — the stationary wavelet transform image transform (``SWTTransform``), and
* ``engine``     — the retrieval evaluator (``CustomCalculator.calculate_maphashing``, ``get_knn``),

behind the reference's own Python plugin API, on hand-written CUDA kernels reached through the C-ABI of
``libb200ret.so`` (``include/b200ret.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _cabi  # noqa: F401  (does not load the library until first use)
