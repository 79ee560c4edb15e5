"""Loads the reference's OWN evaluator modules, unmodified, from ``/root/reference`` (test / baseline infrastructure).

``main/engine/accuracy_calculator.py`` and ``main/engine/get_knn.py`` import third-party packages that are not installed
here (pytorch_metric_learning, torchmetrics, faiss), but the functions on the hot path — ``CustomCalculator.
calculate_maphashing / calc_hamming_dist / label_comparison_fn / per_bit_balance`` and ``get_knn_torch`` — only use torch.
The absent modules are replaced by empty stubs (plus the two PML helpers the calculator calls, restated from
pytorch-metric-learning's published source) and the reference files are executed from where they lie.  Used by
``tests/golden/make_golden_eval.py`` (golden vectors) and by ``bench.py --impl reference`` (CPU baseline of kind
"reference") whenever the reference tree is readable; the GPU box has no ``/root/reference`` and never gets here.
"""
import importlib.util
import logging
import os
import sys
import types

import torch

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "main/engine/accuracy_calculator.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference():
    class AccuracyCalculator:                      # stand-in for PML's base class: ctor args only
        def __init__(self, include=(), exclude=(), avg_of_avgs=False, return_per_class=False, k=None,
                     label_comparison_fn=None, device=None, knn_func=None, kmeans_func=None):
            self.k = k
            self.device = device or torch.device("cpu")

        def requires_knn(self):
            return []

    def get_label_match_counts(query_labels, reference_labels, label_comparison_fn):
        # pytorch-metric-learning's published helper (custom-comparison branch)
        uniq = torch.unique(query_labels, dim=0)
        counts = torch.empty(len(uniq), dtype=torch.long)
        for i in range(len(uniq)):
            counts[i] = torch.sum(label_comparison_fn(uniq[i:i + 1], reference_labels))
        return uniq, counts

    _stub("pytorch_metric_learning")
    _stub("pytorch_metric_learning.utils")
    _stub("pytorch_metric_learning.utils.common_functions", numpy_to_torch=torch.as_tensor)
    _stub("pytorch_metric_learning.utils.accuracy_calculator", AccuracyCalculator=AccuracyCalculator,
          get_label_match_counts=get_label_match_counts, get_lone_query_labels=None)
    _stub("torchmetrics")
    _stub("torchmetrics.retrieval", RetrievalRPrecision=None, RetrievalMAP=None,
          RetrievalPrecisionRecallCurve=None, RetrievalPrecision=None)
    _stub("faiss")
    main = _stub("main")
    main.__path__ = []
    _stub("main.utils", LOGGER=logging.getLogger("RETRIEVAL"))
    eng = _stub("main.engine")
    eng.__path__ = [os.path.join(REF, "main/engine")]
    knn = _load("main.engine.get_knn", os.path.join(REF, "main/engine/get_knn.py"))
    acc = _load("main.engine.accuracy_calculator", os.path.join(REF, "main/engine/accuracy_calculator.py"))
    return acc, knn
