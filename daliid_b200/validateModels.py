"""Drop-in mirror of the reference's ``validateModels.py`` (same class and method names,
argument meaning, return values and printed text), with the hot-path arithmetic moved to
the B200 library.

reference                                   here
------------------------------------------  ------------------------------------------------
validateModels.setParameters   (28-32)      identical
validateModels.validate        (35-58)      features -> one fused library call
                                            (normalise, 1 - q.g, rank, CMC/mAP)
validateModels.calculateMetrics(61-76)      evaluate_rank on a given distmat
validateBRIAR.calculateMetrics (84-105)     top-20 selection kernel, no full argsort
validationManager.getValidator (108-118)    identical
MSMT17_validator               (120-197)    centres + dot-similarity + top-5 kernel

Feature extraction stays the reference's (``getFeatures.extractFeatures``, unchanged per
the scope contract); it is looked up lazily so this module imports without the
reference on the path, and can be replaced through ``feature_extractor``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import metrics


def _reference_extract_features(subset, img_height, img_width, model, batch_size, gpu_index):
    from getFeatures import extractFeatures  # the reference's module, unchanged
    return extractFeatures(subset, img_height, img_width, model, batch_size, gpu_index)


class validateModels:

    #: callable(subset, img_height, img_width, model, batch_size, gpu_index) -> [N, D] tensor
    feature_extractor = staticmethod(_reference_extract_features)
    #: arithmetic of the contraction: "auto" (fp32 class on the tensor cores: f16x3 for the unit
    #: rows of the cosine path), "fp32" (exact SIMT path), "tf32" / "f16" (single pass)
    precision = metrics.DEFAULT_PRECISION
    #: keep extracted features on the GPU when the extractor returns CUDA tensors
    ranks = [1, 5, 10]

    def setParameters(self, img_height, img_width, rerank, gpu_index):
        self.img_height = img_height
        self.img_width = img_width
        self.rerank = rerank
        self.gpu_index = gpu_index

    def validate(self, queries, gallery, model):

        model.eval()
        queries_fvs = self.feature_extractor(queries, self.img_height, self.img_width, model, 500, self.gpu_index)
        gallery_fvs = self.feature_extractor(gallery, self.img_height, self.img_width, model, 500, self.gpu_index)

        # `rerank` (setParameters) is IGNORED, exactly like upstream: the reference keeps its
        # re-ranking block commented out (validateModels.py:49-53), so the same setParameters call
        # must give the same cmc / mAP here.  The hook is available as a separate opt-in,
        # `enable_rerank = True`, which no reference caller sets.
        if getattr(self, "enable_rerank", False):
            print('Applying person re-ranking ...')
            # features are L2-normalised once (validateModels.py:41-42); the three matrices are then
            # computed from the unit rows as they are (normalize=False: no second division).
            # torchreid's "euclidean" is the squared distance.
            qn = metrics.normalize(queries_fvs)
            gn = metrics.normalize(gallery_fvs)
            prec = "tf32c" if self.precision in ("auto", "f16x3", "f16") else self.precision
            distmat = metrics.compute_distance_matrix(qn, gn, "cosine", precision=prec, normalize=False)
            distmat_qq = metrics.compute_distance_matrix(qn, qn, "sqeuclidean", precision=prec, normalize=False)
            distmat_gg = metrics.compute_distance_matrix(gn, gn, "sqeuclidean", precision=prec, normalize=False)
            distmat = metrics.re_ranking(distmat, distmat_qq, distmat_gg)
            del queries_fvs, gallery_fvs, distmat_qq, distmat_gg
            cmc, mAP = self.calculateMetrics(distmat, queries, gallery)
            if not isinstance(distmat, torch.Tensor):
                distmat = torch.from_numpy(distmat)
            return cmc, mAP, distmat

        # normalise -> 1 - q.g -> rank -> CMC/mAP, one library call (validateModels.py:41-69)
        print('Computing CMC and mAP ...')
        cmc, mAP, distmat = metrics.evaluate_features(
            queries_fvs, gallery_fvs, queries[:, 1], gallery[:, 1], queries[:, 2], gallery[:, 2],
            metric="cosine", precision=self.precision, return_distmat=True)
        del queries_fvs, gallery_fvs
        self._report(cmc, mAP)
        if not isinstance(distmat, torch.Tensor):
            distmat = torch.from_numpy(distmat)
        return cmc, mAP, distmat

    def _report(self, cmc, mAP):
        print('** Results **')
        print('mAP: {:.2%}'.format(mAP))
        print('Ranks:')
        for r in self.ranks:
            print('Rank-{:<3}: {:.2%}'.format(r, cmc[r - 1]))

    def calculateMetrics(self, distmat, queries, gallery):

        # compute Ranks
        print('Computing CMC and mAP ...')
        cmc, mAP = metrics.evaluate_rank(distmat, queries[:, 1], gallery[:, 1],
                                         queries[:, 2], gallery[:, 2], use_metric_cuhk03=False)
        self._report(cmc, mAP)
        return cmc, mAP


class validateBRIAR(validateModels):

    def __init__(self):
        super(validateModels, self).__init__()

    def validate(self, queries, gallery, model):
        model.eval()
        queries_fvs = self.feature_extractor(queries, self.img_height, self.img_width, model, 500, self.gpu_index)
        gallery_fvs = self.feature_extractor(gallery, self.img_height, self.img_width, model, 500, self.gpu_index)
        distmat = metrics.compute_distance_matrix(queries_fvs, gallery_fvs, "cosine",
                                                  precision=self.precision)
        del queries_fvs, gallery_fvs
        cmc, mAP = self.calculateMetrics(distmat, queries, gallery)
        if not isinstance(distmat, torch.Tensor):
            distmat = torch.from_numpy(distmat)
        return cmc, mAP, distmat

    def calculateMetrics(self, distmat, queries, gallery):

        nq = queries.shape[0]
        q_ids, g_ids = metrics.canonicalize_labels(queries[:, 1], gallery[:, 1])
        gt = q_ids.reshape(nq, 1)
        cmc = []

        # compute Ranks
        ranks = [1, 5, 10, 20]
        print('Computing CMC and mAP ...')
        _, ranked_idx = metrics.topk_identify(distmat, k=20)  # argsort(distmat)[:, :20]
        if isinstance(ranked_idx, torch.Tensor):
            ranked_idx = ranked_idx.cpu().numpy()
        valid = ranked_idx >= 0  # galleries shorter than 20 are padded with -1
        predicted = np.where(valid, g_ids[np.clip(ranked_idx, 0, None)], -1)

        matching = (gt == predicted) & valid

        print('** Results **')
        print('Ranks:')
        for r in ranks:
            rank_value = np.mean(np.sum(matching[:, :r], axis=1) > 0)
            print('Rank-{:<3}: {:.2%}'.format(r, rank_value))
            cmc.append(rank_value)

        return cmc, 0


class validationManager:

    @staticmethod
    def getValidator(dataset_name):

        if dataset_name == "BRIAR":
            validator = validateBRIAR()
        else:
            validator = validateModels()

        return validator


class MSMT17_validator:
    """Mirror of validateModels.py:120-197 (class centres, similarity, top-5, balanced accuracy)."""

    feature_extractor = staticmethod(_reference_extract_features)

    def __init__(self, train_images, val_images, trainer, dir_to_save):

        self.train_images = train_images
        self.val_images = val_images
        self.img_height = trainer.img_height
        self.img_width = trainer.img_width
        self.gpu_index = trainer.gpu_indexes[0]
        self.model_name = trainer.model_name
        self.version = trainer.version
        self.trainer = trainer
        self.best_accuracy = 0.0
        self.best_iter = 0
        self.dir_to_save = dir_to_save

    def validate(self, pipeline_iter):

        balanced_accuracy_online = self.validate_with_valSet(self.trainer.model_online)
        balanced_accuracy_momentum = self.validate_with_valSet(self.trainer.model_momentum)

        if balanced_accuracy_online > self.best_accuracy or balanced_accuracy_momentum > self.best_accuracy:

            if balanced_accuracy_online > balanced_accuracy_momentum:
                self.best_accuracy = balanced_accuracy_online
            else:
                self.best_accuracy = balanced_accuracy_momentum

            self.best_iter = pipeline_iter

            torch.save(self.trainer.model_online.state_dict(), "%s/model_online_bestACC_%s_%s.h5" % (self.dir_to_save, self.model_name, self.version))
            torch.save(self.trainer.model_momentum.state_dict(), "%s/model_momentum_bestACC_%s_%s.h5" % (self.dir_to_save, self.model_name, self.version))

        print("Best Balanced Accuracy: {:.2%} and best iter: {}".format(self.best_accuracy, self.best_iter))

    def validate_with_valSet(self, model):

        model.eval()
        selected_fvs = self.feature_extractor(self.train_images, self.img_height, self.img_width, model, 500, self.gpu_index)
        val_fvs = self.feature_extractor(self.val_images, self.img_height, self.img_width, model, 500, self.gpu_index)
        return self.balanced_accuracy_from_features(selected_fvs, val_fvs)

    def balanced_accuracy_from_features(self, selected_fvs, val_fvs):
        selected_fvs = torch.as_tensor(metrics.normalize(selected_fvs))
        identities_labels = np.int32(self.train_images[:, 1])
        labels = np.unique(identities_labels)

        # class centres: mean of the normalised features of each identity (validateModels.py:170-177)
        centers = torch.stack([torch.mean(selected_fvs[torch.as_tensor(identities_labels == label,
                                                                        device=selected_fvs.device)], dim=0)
                               for label in labels])

        # S = normalise(val) @ normalise(centers).T ; top-5 by similarity (validateModels.py:159-180)
        S = metrics.compute_distance_matrix(val_fvs, centers, metric="dot", normalize=True)
        _, closest_centers_idxes = metrics.topk_identify(S, k=min(5, len(labels)), largest=True)
        if isinstance(closest_centers_idxes, torch.Tensor):
            closest_centers_idxes = closest_centers_idxes.cpu().numpy()
        closest_centers = labels[closest_centers_idxes]

        true_matches = np.int32(self.val_images[:, 1]) == closest_centers[:, 0]

        identities_labels = np.int32(self.val_images[:, 1])
        labels = np.unique(identities_labels)

        balanced_acc = 0.0
        for label in labels:
            predictions = true_matches[identities_labels == label]
            TPR = np.sum(predictions) / predictions.shape[0]
            balanced_acc += TPR

        balanced_acc = balanced_acc / labels.shape[0]
        print("Balanced Accuracy on Validation Set: {:.3%}".format(balanced_acc))

        return balanced_acc
