"""Device-resident feature sink: the caller side of the hot path.

The reference's ``getFeatures.extractFeatures`` (getFeatures.py:47-71) copies every batch of
backbone outputs to the host (``fv.data.cpu()``, :62) and grows the result with
``torch.cat`` (:64-67, quadratic in the number of batches); the evaluation then has to push
the whole ``[N, D]`` matrix back over PCIe.  Here the rows are written straight into one
pre-sized CUDA tensor, which ``metrics.evaluate_features`` / ``validateModels.validate`` take
as they are: no host round trip, and the call returns as soon as the last batch is enqueued.

The backbone, the dataset class and the DataLoader stay the reference's (unchanged per the
scope contract); only the sink differs.  Values are identical to the reference's: the same
``model(batch.cuda())`` outputs, not rounded or reordered.
"""
from __future__ import annotations

import time

import torch

__all__ = ["DeviceFeatureSink", "extract_features_to_device", "extractFeatures"]


class DeviceFeatureSink:
    """Pre-sized ``[capacity, D]`` fp32 CUDA matrix filled batch by batch (D from the first batch)."""

    def __init__(self, capacity, device):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.buf = None
        self.n = 0

    def append(self, fv):
        fv = fv.detach()
        if fv.dim() != 2:
            fv = fv.reshape(fv.shape[0], -1)
        if self.buf is None:
            self.buf = torch.empty((self.capacity, fv.shape[1]), dtype=torch.float32, device=self.device)
        b = fv.shape[0]
        if self.n + b > self.capacity:
            raise ValueError(f"feature sink overflow: {self.n + b} rows > capacity {self.capacity}")
        self.buf[self.n:self.n + b].copy_(fv, non_blocking=True)  # device-to-device, casts to fp32
        self.n += b

    def result(self):
        if self.buf is None:
            return torch.empty((0, 0), dtype=torch.float32, device=self.device)
        return self.buf[:self.n]


def extract_features_to_device(loader, model, gpu_index=0, capacity=None):
    """The loop of getFeatures.py:56-67 with a device sink: ``[N, D]`` fp32 CUDA tensor."""
    model.eval()
    if capacity is None:
        capacity = len(loader.dataset)
    dev = torch.device(f"cuda:{gpu_index}")
    sink = DeviceFeatureSink(capacity, dev)
    with torch.no_grad():
        for batch in loader:
            if isinstance(batch, (list, tuple)):
                batch = batch[0]
            sink.append(model(batch.cuda(gpu_index, non_blocking=True)))
    return sink.result()


def extractFeatures(subset, img_height, img_width, model, batch_size, gpu_index=0, dataset=None,
                    turbulance_dir_path=None, turb_strength=None):
    """Same signature as the reference's ``extractFeatures`` (getFeatures.py:47); returns the
    features on ``cuda:gpu_index`` instead of the host.  Uses the reference's own ``sample``
    dataset class and DataLoader settings (getFeatures.py:51-52)."""
    from torch.utils.data import DataLoader

    from getFeatures import sample  # the reference's module, unchanged

    data = sample(dataset, subset, turbulance_dir_path, turb_strength, img_height, img_width)
    loader = DataLoader(data, batch_size=batch_size, num_workers=8, pin_memory=True)
    start = time.time()
    fvs = extract_features_to_device(loader, model, gpu_index, capacity=len(data))
    torch.cuda.synchronize(gpu_index)
    end = time.time()
    print("Features extracted in %.2f seconds" % (end - start))
    return fvs
