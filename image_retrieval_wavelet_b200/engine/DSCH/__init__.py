"""Mirror of the evaluation helpers of ``main.engine.DSCH._utils`` (the reference's second Hamming evaluator)."""
from ._utils import (calc_hamming_dist, get_precision_recall_by_Hamming_Radius, mean_average_precision, p_topK, pr_curve)

__all__ = ["calc_hamming_dist", "get_precision_recall_by_Hamming_Radius", "mean_average_precision", "p_topK", "pr_curve"]
