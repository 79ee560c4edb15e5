"""CPU oracle for the two hot paths — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's algorithms (SWT image
transform and Hamming-mAP@k evaluator).  It exists so the CUDA path can be
checked; it is never the thing shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``image_retrieval_wavelet_b200/``
imports it (``tests/test_no_oracle_in_product.py`` enforces that).

Parity status:

* HP-EVAL (``eval_ref``): PINNED.  ``tests/golden/make_golden_eval.py`` runs the
  reference's own ``CustomCalculator`` code (``/root/reference/main/engine/
  accuracy_calculator.py`` with the absent third-party imports stubbed) and
  stores its outputs under ``tests/golden/``; ``tests/test_oracle_eval.py``
  checks this oracle against those.
* HP-SWT (``swt_ref``): PARITY UNPINNED against PyWavelets itself — ``pywt`` is
  not installed, not vendored under ``/root/reference`` and not even pinned in
  its ``requirements.txt``.  The restatement follows PyWavelets' published
  ``swt2`` algorithm and is anchored on the PyWavelets documentation examples
  (``swt([1..8], 'db1', level=2)``) plus analytic known answers.
"""
