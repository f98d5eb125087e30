"""schicexplorer_b200 -- B200-native (sm_100a) MinHash approximate-kNN path of scHicClusterMinHash.

Public surface = the two classes scHiCExplorer imports from ``sparse_neighbors_search``
(/root/reference/schicexplorer/scHicClusterMinHash.py:21-22).
"""
from .minhash import MinHash
from .clustering import MinHashClustering
from .knn import NearestNeighbors

__all__ = ["MinHash", "MinHashClustering", "NearestNeighbors"]
__version__ = "0.1.0"
