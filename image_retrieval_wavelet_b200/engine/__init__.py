"""HP-EVAL: mirrors ``main.engine`` of the reference for the retrieval-evaluator hot path."""
from .accuracy_calculator import AccuracyCalculator, CustomCalculator, get_accuracy_calculator
from .batch_map import build_batch_map_calculator, build_fast_eval_subset, compute_batch_map
from .evaluate import EmbeddingSet, compute_all_embeddings, evaluate, evaluate_multi_k
from .get_knn import get_knn, get_knn_faiss, get_knn_torch
from .make_subset import make_subset
from .map_engine import HammingMapEngine

__all__ = ["AccuracyCalculator", "CustomCalculator", "get_accuracy_calculator", "build_batch_map_calculator",
           "build_fast_eval_subset", "compute_batch_map", "EmbeddingSet", "compute_all_embeddings", "evaluate",
           "evaluate_multi_k", "get_knn", "get_knn_faiss", "get_knn_torch", "make_subset", "HammingMapEngine"]
