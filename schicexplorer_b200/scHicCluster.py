"""``scHicCluster`` with its ``-drm knn`` step on the GPU.

Flag contract of ``/root/reference/schicexplorer/scHicCluster.py:29-124`` (same names, short forms, defaults and
choices); ``main`` follows ``:127-210``: load the cells into one (cells x bins^2) CSR, reduce (``none`` | ``knn`` |
``pca``), cluster (``spectral`` | ``kmeans``), write ``<cell name> <cluster id>`` lines. The one compute-heavy step, the
exact euclidean kNN graph of ``:179-180`` (sklearn ``NearestNeighbors(...).kneighbors_graph(mode='distance')``, brute
force on 7.5 M-dimensional sparse rows), runs on the B200 through :class:`schicexplorer_b200.knn.NearestNeighbors`;
the PCA on top of it (``-pca``, ``:183-188``) through :class:`schicexplorer_b200.pca.PCA`; ``-drm pca``, k-means and
spectral clustering stay on the host (numpy / scikit-learn) exactly as in the reference.
The cell-type colouring / LaTeX-table extras of the reference (``-cct``, ``-ccb``, ``-lt``) are accepted and skipped
with a warning (plot cosmetics, out of scope per SURVEY 8).
"""
from __future__ import annotations

import argparse
import logging

import numpy as np

from . import __version__
from .knn import NearestNeighbors
from .scHicClusterMinHash import load_cells

log = logging.getLogger(__name__)


def parse_arguments(args=None):
    """Returns the parser, like the reference (scHicCluster.py:29)."""
    parser = argparse.ArgumentParser(
        formatter_class=argparse.ArgumentDefaultsHelpFormatter, add_help=False,
        description='scHicCluster uses kmeans or spectral clustering to associate each cell to a cluster. The clustering '
                    'can be run on the raw data, on a kNN graph computed via the exact euclidean distance (on a B200 '
                    'GPU) or via PCA.')
    req = parser.add_argument_group('Required arguments')
    req.add_argument('--matrix', '-m', required=True, metavar='scool scHi-C matrix',
                     help='Single-cell Hi-C matrices to cluster (scool; also: sparse-matrix-files folder, cells npz).')
    req.add_argument('--numberOfClusters', '-c', required=False, default=12, type=int, help='Number of clusters.')
    req.add_argument('--clusterMethod', '-cm', choices=['spectral', 'kmeans'], default='spectral',
                     help='Algorithm to cluster the Hi-C matrices.')
    opt = parser.add_argument_group('Optional arguments')
    opt.add_argument('--chromosomes', nargs='+', help='Chromosomes to include.')
    opt.add_argument('--intraChromosomalContactsOnly', '-ic', action='store_true',
                     help='Load only the intra-chromosomal contacts.')
    opt.add_argument('--additionalPCA', '-pca', action='store_true', help='Compute a PCA on top of the k-nn graph.')
    opt.add_argument('--dimensionsPCA', '-dim_pca', default=20, type=int,
                     help='Number of PCA dimensions considered for clustering.')
    opt.add_argument('--dimensionReductionMethod', '-drm', choices=['none', 'knn', 'pca'], default='none',
                     help='Dimension reduction: knn with euclidean distance, pca, or none.')
    opt.add_argument('--createScatterPlot', '-csp', required=False, default=None,
                     help='Scatter plot of the first two principal components of the matrix clustered on.')
    opt.add_argument('--numberOfNearestNeighbors', '-k', required=False, default=100, type=int,
                     help='Neighbours per cell in the knn graph (at most the number of cells - 1).')
    opt.add_argument('--dpi', '-d', required=False, default=300, type=int, help='Plot resolution.')
    opt.add_argument('--outFileName', '-o', required=True, default='clusters.txt',
                     help='File with one "<cell name> <cluster id>" line per cell.')
    opt.add_argument('--cell_coloring_type', '-cct', required=False, help='Accepted, not used (plot colouring).')
    opt.add_argument('--cell_coloring_batch', '-ccb', required=False, help='Accepted, not used (plot colouring).')
    opt.add_argument('--latexTable', '-lt', help='Accepted, not used (overlap statistics table).')
    opt.add_argument('--figuresize', type=float, nargs=2, default=(15, 6), metavar=('x-size', 'y-size'))
    opt.add_argument('--colorMap', default='glasbey_dark', help='Accepted; matplotlib default colours are used.')
    opt.add_argument('--fontsize', type=float, default=15)
    opt.add_argument('--threads', '-t', required=False, default=8, type=int, help='Host threads (scikit-learn steps).')
    opt.add_argument('--help', '-h', action='help', help='show this help message and exit')
    opt.add_argument('--version', action='version', version='%(prog)s {}'.format(__version__))
    return parser


def knn_graph(X, n_neighbors: int):
    """scHicCluster.py:179-180 on the GPU: exact euclidean kNN graph (n x n CSR, k entries per row, self excluded)."""
    nbrs = NearestNeighbors(n_neighbors=n_neighbors, algorithm='ball_tree').fit(X)
    try:
        return nbrs.kneighbors_graph(mode='distance')
    finally:
        nbrs.close()


def main(args=None):
    args = parse_arguments().parse_args(args)
    from . import loader
    for flag in ('cell_coloring_type', 'cell_coloring_batch', 'latexTable'):
        if getattr(args, flag):
            log.warning('--%s is accepted for compatibility but not used', flag)
    cells, chroms = load_cells(args.matrix, args.threads)
    names = [c.name for c in cells]
    neighborhood_matrix, _ = loader.cells_to_csr(cells, chroms, chromosomes=args.chromosomes,
                                                 intra_only=args.intraChromosomalContactsOnly)

    reduce_to_dimension = neighborhood_matrix.shape[0] - 1
    if args.dimensionReductionMethod == 'knn':
        if args.numberOfNearestNeighbors > reduce_to_dimension:
            args.numberOfNearestNeighbors = reduce_to_dimension
        neighborhood_matrix = knn_graph(neighborhood_matrix, args.numberOfNearestNeighbors)
        if args.additionalPCA:
            from .pca import PCA                     # :183-188 on the GPU; only the kept columns are extracted
            if args.dimensionsPCA:
                args.dimensionsPCA = min(args.dimensionsPCA, neighborhood_matrix.shape[0])
            pca = PCA(n_components=min(neighborhood_matrix.shape) - 1, keep=args.dimensionsPCA or None)
            neighborhood_matrix = pca.fit_transform(neighborhood_matrix)
    elif args.dimensionReductionMethod == 'pca':
        from scipy import linalg
        corrmatrix = np.cov(np.asarray(neighborhood_matrix.todense()))
        _, eigs = linalg.eig(corrmatrix)
        # as the reference (:190-192): the n-1 leading eigenvectors become the samples, so n-1 labels are written
        neighborhood_matrix = np.real(eigs[:, :reduce_to_dimension]).transpose()

    from sklearn.cluster import KMeans, SpectralClustering
    if args.clusterMethod == 'spectral':
        clustering = SpectralClustering(n_clusters=args.numberOfClusters, n_jobs=args.threads,
                                        n_neighbors=reduce_to_dimension, affinity='nearest_neighbors', random_state=0,
                                        eigen_solver='arpack')
    else:
        clustering = KMeans(n_clusters=args.numberOfClusters, random_state=0, n_init=10)
    labels = clustering.fit_predict(neighborhood_matrix)

    if args.createScatterPlot:
        try:
            _scatter_plot(args, neighborhood_matrix, labels)
        except ImportError:
            log.warning('matplotlib is not installed: skipping %s', args.createScatterPlot)

    np.savetxt(args.outFileName, np.array(list(zip(names, labels))), fmt="%s")
    return labels


def _scatter_plot(args, matrix, labels):
    import matplotlib
    matplotlib.use('Agg')
    import matplotlib.pyplot as plt
    from sklearn.decomposition import PCA
    emb = np.asarray(matrix.todense()) if hasattr(matrix, 'todense') else np.asarray(matrix)
    if not (args.dimensionReductionMethod == 'knn' and args.additionalPCA) and args.dimensionReductionMethod != 'pca':
        emb = PCA(n_components=min(min(emb.shape) - 1, 2)).fit_transform(emb)
    plt.figure(figsize=tuple(args.figuresize))
    for lab in sorted(set(labels.tolist())):
        m = labels == lab
        plt.scatter(emb[m, 0], emb[m, 1], s=20, alpha=0.7, label=str(lab))
    plt.legend(bbox_to_anchor=(1.05, 1), loc='upper left', fontsize=args.fontsize)
    plt.xticks([]); plt.yticks([])
    plt.xlabel('PC1', fontsize=args.fontsize); plt.ylabel('PC2', fontsize=args.fontsize)
    name = args.createScatterPlot if '.' in args.createScatterPlot else args.createScatterPlot + '.png'
    plt.tight_layout()
    plt.savefig(name, dpi=args.dpi)
    plt.close()


if __name__ == '__main__':
    main()
