"""Support-set policy (API of the reference's nwhead/support.py) with a device-resident bank.

SupportSetTrain samples (images, labels) for an episodic step on the host exactly like the reference.
SupportSetEval owns the GPU banks built by NWNet.precompute(): the full bank, the cluster-centroid
bank and the per-call random subset, all in the SupportBank layout.
"""
import numpy as np
import torch

from .bank import SupportBank
from .utils import (DatasetMetadata, FeatureDataset, FullDataset, InfiniteUniformClassLoader, KNN,
                    class_centroids)


class SupportSet:
    '''Support set base class for NW (reference nwhead/support.py:7-56).  Environment / IRM splitting
    (env_array, lists of datasets) is outside the accelerated path and is not provided.'''

    def __init__(self, support_set, n_classes, env_array=None):
        if env_array is not None or isinstance(support_set, (list, tuple)):
            raise NotImplementedError(
                'environment-split supports (IRM training, mode="ensemble") are outside the B200 hot path')
        self.y_array = np.array(support_set.targets)
        self.n_classes = n_classes
        self.env_array = np.zeros(len(support_set))
        self.combined_dataset = DatasetMetadata(support_set, self.env_array)
        self.env_map = {0.0: 0}
        env = torch.utils.data.Subset(self.combined_dataset, np.arange(len(support_set)))
        env.targets = self.y_array
        self.env_datasets = [env]


class SupportSetTrain(SupportSet):
    '''Support set for NW training (reference nwhead/support.py:58-93).'''

    def __init__(self, support_set, n_classes, train_type, n_shot, n_way=None, env_array=None):
        super().__init__(support_set, n_classes, env_array)
        if train_type != 'random':
            raise NotImplementedError('train_type="irm" is outside the B200 hot path')
        self.train_type = train_type
        self.n_shot = n_shot
        self.n_way = n_way
        self.train_iter = InfiniteUniformClassLoader(self.combined_dataset, self.n_shot, self.n_way)

    def get_support(self, y):
        '''Samples a support for training: n_way classes that include every query class, n_shot items
        each (host numpy sampling in the reference's call order).'''
        return self.train_iter.next(y)


class SupportSetEval(SupportSet):
    '''Support set for NW evaluation (reference nwhead/support.py:95-165).'''

    def __init__(self, support_set, n_classes, n_shot_random, n_shot_full, n_shot_cluster=3, n_neighbors=20,
                 env_array=None, kernel_type='euclidean', precision='auto'):
        super().__init__(support_set, n_classes, env_array)
        self.n_shot_random = n_shot_random
        self.n_shot_full = n_shot_full
        self.n_shot_cluster = n_shot_cluster
        self.n_neighbors = n_neighbors
        self.kernel_type = kernel_type
        self.precision = precision
        self.support_loaders = self._build_full_loader()

    def build_infer_iters(self, sfeat, sy, smeta, sfeat_env=None, sy_env=None, smeta_env=None):
        '''Builds the device banks for every inference mode (reference nwhead/support.py:113-133).'''
        # Full
        self.full_feat, self.full_y, self.full_meta = sfeat, sy, smeta
        self.full_feat_sep, self.full_y_sep, self.full_meta_sep = sfeat_env, sy_env, smeta_env
        self.full_bank = SupportBank.build(sfeat, sy, self.n_classes, self.kernel_type, self.precision)

        # Cluster: n_shot_cluster == 1 -> class means reduced on the GPU from the fp32 features
        if self.n_shot_cluster != 1:
            raise NotImplementedError("n_shot_cluster > 1 (host-side KMeans in the reference) is outside the B200 path")
        self.cluster_feat, self.cluster_y = class_centroids(
            sfeat if sfeat.stride(1) == 1 else sfeat.contiguous(), self.full_bank.perm, self.full_bank.offsets,
            self.n_classes)
        self.cluster_bank = SupportBank.build(self.cluster_feat, self.cluster_y, self.n_classes, self.kernel_type,
                                              self.precision)

        # Random: host-side index sampling over the bank rows, device-side gather
        self.random_iter = InfiniteUniformClassLoader(FeatureDataset(sfeat, sy.cpu(), smeta), self.n_shot_random)
        inv = None
        if self.full_bank.perm is not None:
            inv = torch.empty_like(self.full_bank.perm)
            inv[self.full_bank.perm] = torch.arange(len(inv), device=inv.device)
        self._source_to_bank_row = inv

        # KNN (exact, on the GPU).  HNSW needs hnswlib and is not provided.
        self.knn = KNN(self.full_feat, self.full_y, n_neighbors=self.n_neighbors)

    def get_support(self, mode, x=None):
        '''Returns the support for an inference mode: a SupportBank for 'full' / 'cluster' / 'random',
        a (features, labels) pair for 'knn'.'''
        try:
            if mode == 'random':
                idx = torch.as_tensor(self.random_iter.sample_indices(), device=self.full_bank.device)
                if self._source_to_bank_row is not None:
                    idx = torch.sort(self._source_to_bank_row[idx]).values
                return self.full_bank.subset(idx)
            elif mode == 'full':
                return self.full_bank
            elif mode == 'cluster':
                return self.cluster_bank
            elif mode == 'knn':
                return self.knn(x)
            elif mode in ('ensemble', 'hnsw'):
                raise NotImplementedError(f'mode={mode!r} is outside the B200 hot path')
            else:
                raise NotImplementedError
        except AttributeError:
            raise AttributeError('Did you run precompute()?')

    def _build_full_loader(self):
        '''Class-balanced loader for precomputing features (reference nwhead/support.py:156-165).'''
        self.full_datasets = [FullDataset(env, self.n_shot_full) for env in self.env_datasets]
        return [torch.utils.data.DataLoader(env, batch_size=128, shuffle=False, num_workers=0)
                for env in self.full_datasets]
