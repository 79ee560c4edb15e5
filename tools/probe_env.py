"""What the GPU box offers beyond this container: PyWavelets? peer access? (run under gpurun)"""
import importlib
import torch

for mod in ("pywt", "pytorch_wavelets", "faiss", "pytorch_metric_learning", "torchmetrics"):
    try:
        m = importlib.import_module(mod)
        print(mod, "present", getattr(m, "__version__", "?"))
    except Exception as e:
        print(mod, "absent:", type(e).__name__)
n = torch.cuda.device_count()
print("gpus", n, torch.cuda.get_device_name(0))
for i in range(n):
    for j in range(n):
        if i != j:
            print("p2p", i, j, torch.cuda.can_device_access_peer(i, j))
