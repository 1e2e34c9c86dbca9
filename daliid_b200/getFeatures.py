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

__all__ = ["DeviceFeatureSink", "extract_features_to_device", "extractFeatures", "rank_by_similarity",
           "get_subset", "get_subset_one_encoder"]


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


def rank_by_similarity(selected_fvs, train_fvs):
    """Training-set indices ordered by similarity to ONE selected sample, best first -- the
    arithmetic of ``get_subset`` / ``get_subset_one_encoder`` (getFeatures.py:243-353) once the
    features exist.  ``selected_fvs`` / ``train_fvs``: one ``[1,D]`` / ``[N,D]`` pair per encoder
    (a list of up to three, or single matrices).  Per encoder: L2-normalise both sides
    (:265-267, :285-287), ``sim = sel @ train.T`` (:289-291); several encoders are averaged in the
    reference's order ``(sim01 + sim02 + sim03)/3`` (:293); the result is
    ``torch.argsort(sim, dim=1, descending=True)[0]`` (:303) as an int64 tensor on the features'
    device.  Distances, fusion and the ordering run in the C-ABI library."""
    from . import metrics
    if not isinstance(selected_fvs, (list, tuple)):
        selected_fvs, train_fvs = [selected_fvs], [train_fvs]
    if len(selected_fvs) != len(train_fvs) or not 1 <= len(selected_fvs) <= 3:
        raise ValueError("one selected / training feature pair per encoder, at most three encoders")
    sims = []
    for sel, tr in zip(selected_fvs, train_fvs):
        if sel.shape[0] != 1:
            raise ValueError("selected_fvs holds the features of exactly one sample ([1, D])")
        sims.append(metrics.compute_distance_matrix(sel, tr, "dot", normalize=True, padded=False))
    sim = sims[0] if len(sims) == 1 else metrics.fuse_distmats(sims)
    order = metrics.argsort_rows(sim, descending=True)[0]
    return order.long() if isinstance(order, torch.Tensor) else torch.from_numpy(order).long()


def _selected_image(selected_sample):
    from getFeatures import transform_person, transform_vehicle  # the reference's module, unchanged
    import torchreid
    img = torchreid.utils.tools.read_image(selected_sample[0])
    tf = transform_person if selected_sample[3] == "person" else transform_vehicle
    return torch.stack([tf(img)])


def _train_loader(train_set, batch_size):
    from torch.utils.data import DataLoader

    from getFeatures import sample  # the reference's dataset class
    return DataLoader(sample(train_set), batch_size=batch_size, shuffle=False, num_workers=8, pin_memory=True,
                      drop_last=False)


def get_subset(selected_sample, train_set, perc_closest, encoder01, encoder02, encoder03, batch_size=500,
               gpu_index=0):
    """Same signature and result as getFeatures.py:243-309 (the ``perc_closest`` nearest training
    samples of ``selected_sample`` by mean similarity of three encoders); features stay on the GPU
    and the similarity / ordering run in the library."""
    start = time.time()
    img = _selected_image(selected_sample).cuda(gpu_index)
    sel, train = [], []
    for enc in (encoder01, encoder02, encoder03):
        enc.eval()
        with torch.no_grad():
            sel.append(enc(img).detach().float())
        train.append(extract_features_to_device(_train_loader(train_set, batch_size), enc, gpu_index,
                                                capacity=len(train_set)))
    order = rank_by_similarity(sel, train)
    topK = int(len(train_set) * perc_closest)
    subset = train_set[order[:topK].cpu().numpy()]
    print("Subset calculated in %.2f seconds" % (time.time() - start))
    return subset


def get_subset_one_encoder(selected_sample, train_set, topK, encoder, batch_size=500, gpu_index=0):
    """Same signature and result as getFeatures.py:311-353: ``(selected_indexes,
    non_selected_indexes)`` of the training set by similarity under one encoder."""
    start = time.time()
    img = _selected_image(selected_sample).cuda(gpu_index)
    encoder.eval()
    with torch.no_grad():
        sel = encoder(img).detach().float()
    train = extract_features_to_device(_train_loader(train_set, batch_size), encoder, gpu_index,
                                       capacity=len(train_set))
    order = rank_by_similarity(sel, train).cpu()
    print("Subset calculated in %.2f seconds" % (time.time() - start))
    return order[:topK], order[topK:]
