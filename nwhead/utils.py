"""`nwhead.utils` of the reference (nwhead/utils.py:7-246) -> nwhead_b200.utils."""
from nwhead_b200.utils import (HNSW, KNN, DatasetMetadata, FeatureDataset, FullDataset,  # noqa: F401
                               InfiniteUniformClassLoader, compute_clusters, get_separated_indices)
