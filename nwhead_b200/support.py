"""Support-set policy (API of the reference's nwhead/support.py) with a device-resident bank.

SupportSetTrain samples (images, labels) for an episodic step on the host exactly like the reference.
SupportSetEval owns the GPU banks built by NWNet.precompute(): the full bank, the cluster-centroid
bank and the per-call random subset, all in the SupportBank layout.
"""
import numpy as np
import torch

from .bank import SupportBank
from .utils import (DatasetMetadata, FeatureDataset, FullDataset, HNSW, InfiniteUniformClassLoader, KNN,
                    class_centroids, kmeans_centroids)


class SupportSet:
    '''Support set base class for NW (reference nwhead/support.py:7-56).

    env_array: optional array with one integer environment indicator per support item (IRM training and
    mode='ensemble').  Without it the whole dataset is a single environment.  (The reference's third form, a
    list of datasets, reads `.targets` of the list itself and cannot be constructed; it is not provided.)'''

    def __init__(self, support_set, n_classes, env_array=None):
        if isinstance(support_set, (list, tuple)):
            raise NotImplementedError('pass one dataset plus env_array; lists of datasets are not supported')
        self.y_array = np.array(support_set.targets)
        self.n_classes = n_classes
        self.env_array = np.zeros(len(support_set)) if env_array is None else np.asarray(env_array)
        self.combined_dataset = DatasetMetadata(support_set, self.env_array)
        self.env_datasets = self._separate_env_datasets(self.combined_dataset)

    def _separate_env_datasets(self, combined_dataset):
        '''One Subset per environment value, in sorted order of the values (reference nwhead/support.py:47-56).'''
        env_datasets = []
        self.env_map = {}
        for i, attr in enumerate(np.unique(self.env_array)):
            self.env_map[attr] = i
            indices = (self.env_array == attr).nonzero()[0]
            env = torch.utils.data.Subset(combined_dataset, indices)
            env.targets = self.y_array[indices]
            env_datasets.append(env)
        return env_datasets


class SupportSetTrain(SupportSet):
    '''Support set for NW training (reference nwhead/support.py:58-93).'''

    def __init__(self, support_set, n_classes, train_type, n_shot, n_way=None, env_array=None):
        super().__init__(support_set, n_classes, env_array)
        self.train_type = train_type
        self.n_shot = n_shot
        self.n_way = n_way
        if train_type == 'random':
            self.train_iter = InfiniteUniformClassLoader(self.combined_dataset, self.n_shot, self.n_way)
        else:  # 'irm': one sampler per environment, every class of the environment, n_shot each
            self.train_iter = [iter(InfiniteUniformClassLoader(env, self.n_shot)) for env in self.env_datasets]

    def get_support(self, y):
        '''Samples a support for training (host numpy sampling in the reference's call order).
        'random': n_way classes that include every query class, n_shot items each.
        'irm'   : a random environment, then n_shot items of each of its classes.'''
        if self.train_type == 'irm':
            train_iter = np.random.choice(self.train_iter)
            return train_iter.next()
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
        # one bank per environment for mode='ensemble' (reference nwhead/nw.py:143-154)
        self.env_banks = None
        if sfeat_env is not None and len(sfeat_env) > 1:
            self.env_banks = [SupportBank.build(f, y, self.n_classes, self.kernel_type, self.precision)
                              for f, y in zip(sfeat_env, sy_env)]

        # Cluster: n_shot_cluster == 1 -> class means reduced on the GPU from the fp32 features;
        # n_shot_cluster > 1 -> per-class k-means for all classes at once (nw_kmeans_assign + nw_class_centroids)
        cfeat = sfeat if sfeat.stride(1) == 1 else sfeat.contiguous()
        if self.n_shot_cluster == 1:
            self.cluster_feat, self.cluster_y = class_centroids(cfeat, self.full_bank.perm, self.full_bank.offsets,
                                                                self.n_classes)
        else:
            self.cluster_feat, self.cluster_y = kmeans_centroids(
                cfeat, sy.to(torch.int32).contiguous(), self.full_bank.perm, self.full_bank.offsets, self.n_classes,
                self.n_shot_cluster)
        self.cluster_bank = SupportBank.build(self.cluster_feat, self.cluster_y, self.n_classes, self.kernel_type,
                                              self.precision)

        # Random: host-side index sampling over the bank rows, device-side gather
        self.random_iter = InfiniteUniformClassLoader(FeatureDataset(sfeat, sy.cpu(), smeta), self.n_shot_random)
        inv = None
        if self.full_bank.perm is not None:
            inv = torch.empty_like(self.full_bank.perm)
            inv[self.full_bank.perm] = torch.arange(len(inv), device=inv.device)
        self._source_to_bank_row = inv

        # KNN and "HNSW": both exact on the GPU (no hnswlib index to build)
        self.knn = KNN(self.full_feat, self.full_y, n_neighbors=self.n_neighbors, bank=self.full_bank)
        self.hnsw = HNSW(self.full_feat, self.full_y, n_neighbors=self.n_neighbors, bank=self.full_bank)

    def get_support(self, mode, x=None, raw=False):
        '''Returns the support for an inference mode: a SupportBank for 'full' / 'cluster' / 'random', a list of
        per-environment SupportBanks for 'ensemble'; 'knn' / 'hnsw': a SupportBank over the neighbours (a (features,
        labels) pair when they are fewer than 26).
        raw=True returns the fp32 (features, labels) tensors the banks were built from instead (differentiable
        predict: the direct path needs the unrounded rows).'''
        try:
            if mode == 'random':
                idx = torch.as_tensor(self.random_iter.sample_indices(), device=self.full_bank.device)
                if raw:
                    return self.full_feat[idx], self.full_y[idx]
                if self._source_to_bank_row is not None:
                    idx = torch.sort(self._source_to_bank_row[idx]).values
                return self.full_bank.subset(idx)
            elif mode == 'full':
                return (self.full_feat, self.full_y) if raw else self.full_bank
            elif mode == 'cluster':
                return (self.cluster_feat, self.cluster_y) if raw else self.cluster_bank
            elif mode == 'knn':
                return self.knn(x, None if raw else (self.n_classes, self.kernel_type, self.precision))
            elif mode == 'ensemble':
                if raw:
                    if self.env_banks is None:
                        return [(self.full_feat, self.full_y)]
                    return list(zip(self.full_feat_sep, self.full_y_sep))
                return self.env_banks if self.env_banks is not None else [self.full_bank]
            elif mode == 'hnsw':
                return self.hnsw(x, None if raw else (self.n_classes, self.kernel_type, self.precision))
            else:
                raise NotImplementedError
        except AttributeError:
            raise AttributeError('Did you run precompute()?')

    def _build_full_loader(self):
        '''Class-balanced loader for precomputing features (reference nwhead/support.py:156-165).'''
        self.full_datasets = [FullDataset(env, self.n_shot_full) for env in self.env_datasets]
        return [torch.utils.data.DataLoader(env, batch_size=128, shuffle=False, num_workers=0)
                for env in self.full_datasets]
