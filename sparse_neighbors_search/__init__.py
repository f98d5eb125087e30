"""Import shim: lets ``from sparse_neighbors_search import MinHash, MinHashClustering``
(/root/reference/schicexplorer/scHicClusterMinHash.py:21-22) resolve to the B200 path unchanged."""
from schicexplorer_b200 import MinHash, MinHashClustering

__all__ = ["MinHash", "MinHashClustering"]
