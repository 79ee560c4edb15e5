"""HP-EVAL: mirrors ``main.engine`` of the reference for the retrieval-evaluator hot path."""
from .accuracy_calculator import AccuracyCalculator, CustomCalculator, get_accuracy_calculator
from .batch_map import build_batch_map_calculator, compute_batch_map
from .get_knn import get_knn, get_knn_faiss, get_knn_torch

__all__ = ["AccuracyCalculator", "CustomCalculator", "get_accuracy_calculator", "build_batch_map_calculator",
           "compute_batch_map", "get_knn", "get_knn_faiss", "get_knn_torch"]
