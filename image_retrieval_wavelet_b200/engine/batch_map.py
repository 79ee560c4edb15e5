"""Mirror of ``/root/reference/main/engine/batch_map.py:9-36`` (in-training self-retrieval mAP proxy).

In training mode the reference feeds *raw logits* here (``multi_dino_attention.py:750`` only applies ``sign`` in eval
mode), which makes its "Hamming distance" a real-valued proxy.  A bit-packed evaluator cannot represent that, so this
mirror binarises explicitly (``on_nonbinary='sign'``: bit = logit > 0) — i.e. it reports the mAP the codes would have
after ``sign()``.  That is a documented deviation, not a silent one.
"""
import torch

from .accuracy_calculator import CustomCalculator


def build_batch_map_calculator(distance_metric, device):
    metric_name = "maphashing" if distance_metric == "hamming" else "map"
    calculator = CustomCalculator(exclude=["NMI", "AMI"], k="max_bin_count", with_faiss=False,
                                  distance_metric=distance_metric, device=device, on_nonbinary="sign")
    return calculator, metric_name


def compute_batch_map(calculator, metric_name, embeddings, labels):
    with torch.no_grad():
        embeddings = embeddings.detach()
        if labels.ndim == 2 and labels.size(1) == 1:
            labels = labels.view(-1)
        result = calculator.get_accuracy(query=embeddings, query_labels=labels, reference=embeddings,
                                         reference_labels=labels, embeddings_come_from_same_source=True,
                                         include=[metric_name])
    return result[metric_name]
